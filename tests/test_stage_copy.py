"""CPU: the index arithmetic of stroke_emit_k's staging area (vkvg_b200/csrc/stroke.cu).  A block's items own the global vertex range
[v0, v1) and index range [x0, x1); slot 0 of a staging array stands for the 16-byte aligned global element at or below the range's first one
(voff = v0 & ~1 for 8-byte vertices, ioff = x0 & ~3 for 4-byte indices), items write slot (global - off), and the copy-out moves whole
16-byte vectors where a vector lies inside the range and single elements at its head and tail.  Modelled here element by element: every
global element of the range is written exactly once, from the slot that holds it, nothing outside the range is touched, and every vector
store is 16-byte aligned on both sides."""
import random

SE_BLOCK, SE_VERTS, SE_INDS = 128, 1536, 4608


def copy_out(lo, hi, per_vec, cap):
    """returns {global index: slot} as the kernel's copy loop writes them, and the list of vector stores (global index of their first element)"""
    off = lo & ~(per_vec - 1)
    assert hi - off <= cap
    written, vectors = {}, []
    for tid in range(SE_BLOCK):
        q = tid
        while per_vec * q < hi - off:
            g = off + per_vec * q
            if g >= lo and g + per_vec <= hi:
                vectors.append(g)
                for e in range(per_vec):
                    assert g + e not in written
                    written[g + e] = per_vec * q + e
            else:
                for e in range(per_vec):
                    if lo <= g + e < hi:
                        assert g + e not in written
                        written[g + e] = per_vec * q + e
            q += SE_BLOCK
    return off, written, vectors


def check(lo, hi, per_vec, cap):
    off, written, vectors = copy_out(lo, hi, per_vec, cap)
    assert sorted(written) == list(range(lo, hi))                 # the whole range, once, nothing else
    assert all(slot == g - off for g, slot in written.items())    # from the slot the items wrote it to
    assert all(g % per_vec == 0 for g in vectors)                 # aligned in global memory (the arrays are 256-byte aligned) ...
    assert all((g - off) % per_vec == 0 for g in vectors)         # ... and in the staging area
    if hi - lo >= 2 * per_vec:
        assert len(vectors) >= (hi - lo) // per_vec - 1           # all but a head and a tail go out as vectors


def test_vertex_and_index_ranges():
    r = random.Random(1)
    for _ in range(400):
        v0 = r.randrange(0, 1 << 20)
        n = r.choice([0, 1, 2, 3, 5, 127, 128, 129, 384, 385, 700, SE_VERTS - 2, SE_VERTS - 1])
        if (v0 & 1) + n <= SE_VERTS:
            check(v0, v0 + n, 2, SE_VERTS)
        x0 = r.randrange(0, 1 << 22)
        m = r.choice([0, 3, 6, 9, 18, 381, 384, 2304, SE_INDS - 6, SE_INDS - 3])
        if (x0 & 3) + m <= SE_INDS:
            check(x0, x0 + m, 4, SE_INDS)
