"""Generates the golden fixtures in this directory from the pure-Python big-integer oracle
(oracle/pyref.py: textbook affine formulas + Python ints, no code shared with the C oracle or the
CUDA engine).  The reference holds no literal vectors; the one deterministic test it
has -- kzg/src/commitment.rs:36-54 -- is reproduced as `kzg_kat`.

    python tests/golden/make_golden.py        # rewrites tests/golden/*.json (deterministic)

All integers are canonical (non-Montgomery) hex strings; points are [x, y] or null for infinity.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyref as o  # noqa: E402


def hx(v):
    return hex(v)


def pt(p):
    return None if p is None else [hx(p[0]), hx(p[1])]


def main():
    o.selfcheck()
    out = {}
    # --- reference KAT: kzg/src/commitment.rs:36-54 --------------------------------------------
    srs = o.srs_from_secret(2, 10)
    w, y = o.kzg_open([1, 2, 3], 1, srs)
    out["kzg_kat"] = {
        "secret": hx(2), "circuit_size": 10, "srs": [pt(p) for p in srs], "poly": [hx(1), hx(2), hx(3)],
        "commitment": pt(o.msm_evaluate_in_s([1, 2, 3], srs)), "open_at": hx(1), "evaluation": hx(y), "witness": pt(w),
    }
    # --- kzg/examples/example.rs shape with a seeded secret --------------------------------------
    rng = o.SplitMix64(0xB200)
    secret = rng.fr()
    srs2 = o.srs_from_secret(secret, 10)
    poly = [5, 3, 0, 1]
    w, y = o.kzg_open(poly, 4, srs2)
    out["kzg_example"] = {
        "secret": hx(secret), "circuit_size": 10, "srs": [pt(p) for p in srs2], "poly": [hx(c) for c in poly],
        "commitment": pt(o.msm_evaluate_in_s(poly, srs2)), "open_at": hx(4), "evaluation": hx(y), "witness": pt(w),
    }
    # --- seeded MSMs with corner cases -----------------------------------------------------------
    msms = []
    for n, seed in ((1, 11), (16, 12), (64, 13)):
        r = o.SplitMix64(seed)
        pts = [o.g1_mul(o.G1, r.fr()) for _ in range(n)]
        sc = [r.fr() for _ in range(n)]
        if n >= 16:
            sc[1] = 0
            sc[2] = 1
            sc[3] = o.R - 1
            pts[4] = None                      # infinity base
            pts[6], sc[6] = pts[5], sc[5]      # repeated term (P + P inside a bucket)
            pts[8], sc[8] = o.g1_neg(pts[7]), sc[7]  # cancelling term (P + -P)
        msms.append({"scalars": [hx(s) for s in sc], "bases": [pt(p) for p in pts],
                     "result": pt(o.msm_evaluate_in_s(sc, pts))})
    # all-zero scalars -> identity (nova/src/r1cs/mod.rs:52-59 use), and zip truncation (scheme.rs:88-91)
    r = o.SplitMix64(14)
    pts = [o.g1_mul(o.G1, r.fr()) for _ in range(8)]
    msms.append({"scalars": [hx(0)] * 8, "bases": [pt(p) for p in pts], "result": None})
    sc = [r.fr() for _ in range(5)]
    msms.append({"scalars": [hx(s) for s in sc], "bases": [pt(p) for p in pts],
                 "result": pt(o.msm_evaluate_in_s(sc, pts))})
    out["msm"] = msms
    with open(os.path.join(HERE, "kzg_msm.json"), "w") as f:
        json.dump(out, f, indent=0)
    # --- NTT ---------------------------------------------------------------------------------------
    ntts = []
    for log_n, seed in ((0, 20), (1, 21), (3, 22), (6, 23), (8, 24)):
        r = o.SplitMix64(seed)
        v = [r.fr() for _ in range(1 << log_n)]
        ntts.append({
            "log_n": log_n, "input": [hx(x) for x in v],
            "fft": [hx(x) for x in o.ntt(v, log_n)],
            "ifft": [hx(x) for x in o.intt(v, log_n)],
            "coset": hx(7),
            "coset_fft": [hx(x) for x in o.ntt(v, log_n, coset=7)],
            "coset_ifft": [hx(x) for x in o.intt(v, log_n, coset=7)],
        })
    r = o.SplitMix64(30)
    a = [r.fr() for _ in range(19)]
    b = [r.fr() for _ in range(46)]
    prod = {"a": [hx(x) for x in a], "b": [hx(x) for x in b], "product": [hx(x) for x in o.poly_mul(a, b)]}
    # short input is zero-padded to the domain (plonk/src/circuit.rs:131-133 dummy gates are skipped)
    short = [r.fr() for _ in range(5)]
    pad = {"log_n": 3, "input": [hx(x) for x in short], "ifft": [hx(x) for x in o.intt(short, 3)]}
    with open(os.path.join(HERE, "ntt.json"), "w") as f:
        json.dump({"omega_2_32": hx(o.ROOT_2_32), "ntt": ntts, "poly_mul": prod, "padded_interpolate": pad}, f, indent=0)
    print("golden fixtures written")


if __name__ == "__main__":
    main()
