"""Replays the symmetric-half spot check of the upload path (sample_symmetric in fermi.jl_b200/csrc/fpt_api_upload.inl: the same
generator, 512 samples, |x - y| <= 1e-12 (|x| + |y|) + 1e-300) on the real-molecule arrays of tests/golden/ -- CPU only, no library call.
It shows which arrays the library would read in full although they are symmetric to rounding ("Limitation"), and
what an absolute floor of 1e-12 of the largest sampled magnitude would decide instead.
    python tools/cpu_symcheck_replay.py  > profiles/r02f_symcheck_replay.txt"""
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MASK = (1 << 64) - 1


def pairs(A, kind):
    """the 512 (x, y) pairs the library compares; kind 0: A[i,a,b,c] / A[i,a,c,b], 1: A[i,j,a,b] / A[j,i,b,a], 2: A[i,a,j,b] / A[j,b,i,a]"""
    n0, n1, n2, n3 = A.shape
    flat = np.asfortranarray(A).ravel(order="F")
    s = 0x9E3779B97F4A7C15

    def nxt(m):
        nonlocal s
        s = (s * 6364136223846793005 + 1442695040888963407) & MASK
        return (s >> 33) % m

    out = []
    for _ in range(512):
        i0, i1, i2, i3 = nxt(n0), nxt(n1), nxt(n2), nxt(n3)
        x = flat[i0 + n0 * (i1 + n1 * (i2 + n2 * i3))]
        if kind == 0:
            y = flat[i0 + n0 * (i1 + n1 * (i3 + n2 * i2))]
        elif kind == 1:
            y = flat[i1 + n0 * (i0 + n1 * (i3 + n2 * i2))]
        else:
            y = flat[i2 + n0 * (i3 + n1 * (i0 + n2 * i1))]
        out.append((x, y))
    return out


def library_check(pp):
    for t, (x, y) in enumerate(pp):
        if abs(x - y) > 1e-12 * (abs(x) + abs(y)) + 1e-300:
            return f"rejected at sample {t}: {x:.6e} vs {y:.6e}"
    return "accepted"


def scaled_floor_check(pp):
    scale = max(max(abs(x), abs(y)) for x, y in pp)
    bad = [(x, y) for x, y in pp if abs(x - y) > 1e-12 * (abs(x) + abs(y)) + 1e-12 * scale]
    return "accepted" if not bad else f"rejected ({len(bad)} pairs)"


if __name__ == "__main__":
    print(f"{'case':22s} {'array':5s} {'MB':>6s}  {'true max |A - A^sym| / max|A|':>30s}  library check / with a floor of 1e-12 max|sample|")
    for name in ("water_631g", "glycine_sto3g", "formaldehyde_631gs", "ammonia_augccpvdz", "water_ccpvtz"):
        g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        arrays = [("T2", g["T2"], 1, (1, 0, 3, 2)), ("OVOV", g["OVOV"], 2, (2, 3, 0, 1))]
        if "OVVV" in g.files:          # the packed fixtures rebuild OVVV from one half: exactly symmetric by construction, not informative
            arrays.insert(0, ("OVVV", g["OVVV"], 0, (0, 1, 3, 2)))
        for label, A, kind, perm in arrays:
            asym = float(np.abs(A - A.transpose(perm)).max() / np.abs(A).max())
            pp = pairs(A, kind)
            print(f"{name:22s} {label:5s} {A.nbytes / 2**20:6.2f}  {asym:30.2e}  {library_check(pp)} / {scaled_floor_check(pp)}")
    print("\n(the library applies the check only to pageable arrays of 4 MB and more; these fixtures are smaller, so their GPU tests are not\n"
          " affected -- the point is what the check decides on arrays of this kind)")
