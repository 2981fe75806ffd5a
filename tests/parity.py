"""Helpers shared by the config-scale parity tests and bench.py's `parity` record: pixel statistics against the bar of
BASELINE.json (<= 1/255 per channel at the 99.9th percentile) and an order-preserving alignment of two vertex streams that
agree except at a few knife-edge sites (a dash boundary that the reference's float32 phase puts on the neighbouring segment,
an arc that gains or loses one vertex)."""
import numpy as np


def pixel_stats(a, b):
    """a, b: (H, W, 4) uint8.  Returns max |a - b|, its 99.9th percentile over channels of all pixels, the number of differing
    pixels and their fraction."""
    assert a.shape == b.shape, (a.shape, b.shape)
    d = np.abs(a.astype(np.int16) - b.astype(np.int16))
    n_diff = int((d.max(axis=-1) > 0).sum())
    # percentile over every channel value: with n_diff / n below 1e-3 / 4 it is 0 without sorting 67M values
    flat = d.ravel()
    nz = flat[flat > 0]
    k = int(np.ceil(0.999 * flat.size))          # rank of the 99.9th percentile
    zeros = flat.size - nz.size
    p999 = 0 if k <= zeros else int(np.sort(nz)[k - zeros - 1])
    return {"max_diff": int(d.max()) if d.size else 0, "p99_9": p999, "n_diff": n_diff, "frac_diff": n_diff / max(1, a.shape[0] * a.shape[1])}


def align_vertices(a, b, tol=2e-3, window=96, probe=6, max_chunk=1 << 20):
    """Walk two (n, 2) float arrays that list the same vertices in the same order except at isolated sites where either holds a
    few extra ones.  Returns dict(matched, max_drift, sites=[(i, j, skipped_a, skipped_b)], unmatched_a, unmatched_b, ok)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    i = j = 0
    matched, drift = 0, 0.0
    sites = []
    ua = ub = 0
    ok = True
    chunk = 1024   # grows while the streams agree, restarts small after a site (thousands of sites must not cost a full-length compare each)
    while i < len(a) and j < len(b):
        n = min(len(a) - i, len(b) - j, chunk)
        d = np.abs(a[i:i + n] - b[j:j + n]).max(axis=1)
        bad = np.nonzero(d > tol)[0]
        k = int(bad[0]) if len(bad) else n
        if k:
            matched += k
            drift = max(drift, float(d[:k].max()))
            i += k
            j += k
        if k == n:
            chunk = min(chunk * 4, max_chunk)
            continue
        chunk = 1024
        # re-synchronise: the smallest (da, db) after which `probe` consecutive vertices agree again
        best = None
        for tot in range(1, 2 * window):
            for da in range(max(0, tot - window), min(tot, window) + 1):
                db = tot - da
                m = min(probe, len(a) - i - da, len(b) - j - db)
                if m <= 0:
                    continue
                if np.abs(a[i + da:i + da + m] - b[j + db:j + db + m]).max() <= tol:
                    best = (da, db)
                    break
            if best:
                break
        if not best:
            ok = False
            break
        sites.append((i, j, best[0], best[1]))
        ua += best[0]
        ub += best[1]
        i += best[0]
        j += best[1]
    tail_a, tail_b = len(a) - i, len(b) - j
    return {"matched": matched, "max_drift": drift, "sites": sites, "unmatched_a": ua + (tail_a if ok else 0), "unmatched_b": ub + (tail_b if ok else 0),
            "ok": ok, "stopped_at": (i, j)}
