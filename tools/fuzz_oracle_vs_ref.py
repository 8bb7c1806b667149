#!/usr/bin/env python3
"""Randomised campaign of the CPU restatement (oracle/bp_oracle.c) against the UNMODIFIED reference build (oracle/_ref),
beyond the committed golden vectors and the seeded cases of tests/test_oracle_vs_ref.py. Build container only
(needs oracle/_ref); minutes of CPU, not part of the test suite.

  fuzz_oracle_vs_ref.py bp  SECONDS   flooding sum-product + floating min-sum on the small N=120 code
  fuzz_oracle_vs_ref.py sw  SECONDS   sliding-window BP on the SC-LDPC code (window 1..14, code_type 0 / 1)
The reference object holds one .pchk per process, hence one mode per run. Every case compares n, the success flag, the
decoded bits, the syndrome, the posteriors (bp) and ALL final e->pr / e->lr messages bit for bit.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_sc_pchk  # noqa: E402
import oraclelib as ol  # noqa: E402


def ratios(rs, N, eps, kind):
    """BSC ratios / soft LLRs with erasures / saturated ratios (inf, 0) / arbitrary non-negative ratios incl. 0"""
    if kind == 0:
        p = max(eps, 0.01)
        return np.where(rs.rand(N) < eps, p / (1 - p), (1 - p) / p)
    if kind == 1:
        llr = np.where(rs.rand(N) < eps, -1.0, 1.0) * rs.uniform(0.2, 6.0, N)
        llr[rs.rand(N) < 0.05] = 0.0
        return np.exp(llr)
    if kind == 2:
        with np.errstate(over="ignore"):
            return np.exp(np.where(rs.rand(N) < eps, -1.0, 1.0) * rs.choice([0.5, 20.0, 46.7, 300.0, 800.0], size=N))
    lr = rs.uniform(0, 3, N)
    lr[rs.rand(N) < 0.05] = 0.0
    return lr


def same(a, b, keys):
    return all(np.array_equal(np.asarray(a[k]).view(np.uint64) if np.asarray(a[k]).dtype == np.float64 else a[k],
                              np.asarray(b[k]).view(np.uint64) if np.asarray(b[k]).dtype == np.float64 else b[k]) for k in keys)


def main():
    mode, seconds = sys.argv[1], float(sys.argv[2])
    rs = np.random.RandomState(2026)
    t0, n = time.time(), 0
    if mode == "bp":
        path = os.path.join(ol.GOLDEN, "small_n120_m60.pchk")
        ref, orc = ol.RefLib(path), ol.Oracle(path)
        while time.time() - t0 < seconds:
            lr = ratios(rs, ref.N, rs.choice([0.0, 0.02, 0.05, 0.08, 0.12, 0.2]), rs.randint(4))
            mi = int(rs.choice([0, 1, 3, 10, 50, 200]))
            a, b = ref.decode(lr, mi, True, True), orc.decode(lr, mi, True, True)
            assert a["n"] == b["n"] and a["ok"] == b["ok"] and same(a, b, ("dblk", "pchk", "post", "pr", "lr")), n
            llr = np.log(np.maximum(lr, 1e-300))
            llr = np.where(np.isfinite(llr), llr, 700.0)
            a, b = ref.decode_minsum(llr, mi), orc.decode_minsum(llr, mi)
            assert a["n"] == b["n"] and a["ok"] == b["ok"] and same(a, b, ("dblk",) + (("post",) if a["n"] else ())), n
            n += 1
    else:
        path = os.path.join(ol.GOLDEN, "sc_z32_l12.pchk")
        _, N, _, _, Mv, Mc = gen_sc_pchk.gen_sc(32, 12, 11)
        ref, orc = ol.RefLib(path), ol.Oracle(path)
        while time.time() - t0 < seconds:
            ct = int(rs.randint(0, 2))
            D = 12 + 3 - 1 if ct == 0 else 12 + (3 - 1) // 2
            win, mi = min(int(rs.randint(1, 15)), D), int(rs.choice([0, 1, 2, 5, 12, 40]))
            lr = ratios(rs, N, rs.choice([0.0, 0.02, 0.04, 0.06, 0.08, 0.1, 0.15]), rs.randint(3))
            a = ref.decode_sw(lr, mi, 12, 3, win, Mv[:D], Mc[:D], code_type=ct, want_msgs=True)
            b = orc.decode_sw(lr, mi, 12, 3, win, Mv[:D], Mc[:D], code_type=ct, want_msgs=True)
            assert a["n"] == b["n"] and a["ok"] == b["ok"] and same(a, b, ("dblk", "pchk", "pr", "lr")), (n, win, mi, ct)
            n += 1
    print("%s: %d random cases identical" % (mode, n))


if __name__ == "__main__":
    main()
