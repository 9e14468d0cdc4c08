"""Generates tests/golden/mas_cases.npz from the REFERENCE's own maximum_path_c (compiled by oracle/build_ref_mas.py into
oracle/_ref/).  Small cases are stored whole; large ones as (seed, shape, lengths) + the token index of every frame, which is
the path in compressed form.  Run here (needs /root/reference); the fixture travels, the reference does not.

    python oracle/build_ref_mas.py && python oracle/make_golden_mas.py
"""
import glob
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def load_ref():
    so = glob.glob(os.path.join(HERE, "_ref", "ref_mas_core*.so"))
    if not so:
        raise SystemExit("run oracle/build_ref_mas.py first")
    spec = importlib.util.spec_from_file_location("ref_mas_core", so[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ref_path(ref, neg_cent, t_ys, t_xs):
    v = np.array(neg_cent, np.float32, copy=True)
    p = np.zeros(v.shape, np.int32)
    ref.maximum_path_c(p, v, np.asarray(t_ys, np.int32), np.asarray(t_xs, np.int32))
    return p


def large_case(seed, B, Ty, Tx):
    """Inputs of a large case are regenerated from the seed on both sides (tests and here)."""
    rng = np.random.RandomState(seed)
    nc = (rng.randn(B, Ty, Tx) * 4).astype(np.float32)
    t_ys = rng.randint(Ty // 2, Ty + 1, size=B).astype(np.int32)
    t_xs = np.array([rng.randint(max(1, Tx // 3), min(Tx, ty) + 1) for ty in t_ys], np.int32)
    return nc, t_ys, t_xs


LARGE = [(101, 4, 400, 100), (102, 2, 1000, 257), (103, 8, 162, 128), (104, 1, 2048, 600)]


def main():
    ref = load_ref()
    out = {}
    rng = np.random.RandomState(7)
    n = 0
    for (B, Ty, Tx) in [(1, 1, 1), (1, 5, 1), (1, 5, 5), (2, 7, 3), (3, 12, 6), (2, 33, 32), (1, 40, 17)]:
        nc = (rng.randn(B, Ty, Tx) * 2).astype(np.float32)
        t_ys = rng.randint(max(1, Ty // 2), Ty + 1, size=B).astype(np.int32)
        t_xs = np.array([rng.randint(1, min(Tx, ty) + 1) for ty in t_ys], np.int32)
        if (B, Ty, Tx) == (1, 5, 5):
            t_ys[:], t_xs[:] = 5, 5           # diagonal: every frame its own token
        out["s%d_nc" % n], out["s%d_ty" % n], out["s%d_tx" % n] = nc, t_ys, t_xs
        out["s%d_path" % n] = ref_path(ref, nc, t_ys, t_xs)
        n += 1
    # ties: quantised scores make value[y-1, x] == value[y-1, x-1] frequent; `<` keeps the index on a tie (core.pyx:33)
    nc = rng.randint(-2, 3, size=(3, 30, 9)).astype(np.float32)
    t_ys, t_xs = np.array([30, 21, 9], np.int32), np.array([9, 9, 9], np.int32)
    out["s%d_nc" % n], out["s%d_ty" % n], out["s%d_tx" % n] = nc, t_ys, t_xs
    out["s%d_path" % n] = ref_path(ref, nc, t_ys, t_xs)
    n += 1
    out["n_small"] = np.array(n)
    for i, (seed, B, Ty, Tx) in enumerate(LARGE):
        nc, t_ys, t_xs = large_case(seed, B, Ty, Tx)
        p = ref_path(ref, nc, t_ys, t_xs)
        out["l%d_meta" % i] = np.array([seed, B, Ty, Tx], np.int64)
        out["l%d_token_of_frame" % i] = np.where(p.sum(2) > 0, p.argmax(2), -1).astype(np.int16)   # -1 beyond t_y
        assert (p.sum(2) <= 1).all()
    np.savez_compressed(os.path.join(HERE, "..", "tests", "golden", "mas_cases.npz"), **out)
    print("wrote tests/golden/mas_cases.npz:", n, "small cases,", len(LARGE), "large")


if __name__ == "__main__":
    main()
