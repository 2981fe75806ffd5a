"""TEST FIXTURE GENERATOR (run in the build container, where /root/reference exists):

    python -m tests.golden.make_svg_golden

Copies the reference's SVG test documents (tests/data/*.svg: data fixtures SURVEY.md §8c lists for reuse) next to the
documents authored for this repository (extra_*.svg) under tests/golden/svg/, and dumps each of them with the
reference's own nanoSVG (oracle/_ref/nsvg_dump, built from /root/reference/src/nsvg/nanosvg.h by oracle/Makefile) into
<name>.nsvg.bin — the flat shape list vkvg_svg_render walks.  tests/test_svg.py compares vkvg_b200's parser with them."""
import glob
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("VKVG_REF", "/root/reference")


def main():
    out = os.path.join(HERE, "svg")
    os.makedirs(out, exist_ok=True)
    for f in sorted(glob.glob(os.path.join(REF, "tests", "data", "*.svg"))):
        dst = os.path.join(out, os.path.basename(f))
        shutil.copyfile(f, dst)
        os.chmod(dst, 0o644)
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "nsvg"], check=True)
    tool = os.path.join(ROOT, "oracle", "_ref", "nsvg_dump")
    for f in sorted(glob.glob(os.path.join(out, "*.svg"))):
        subprocess.run([tool, f, f[:-4] + ".nsvg.bin"], check=True)


if __name__ == "__main__":
    main()
