#!/usr/bin/env python3
"""Golden vectors for the sliding-window decoder from the UNMODIFIED reference's Run_SW_Decoder
(dec.cpp:2092-2196) through oracle/_ref (run in the build container only; outputs are committed).

  tests/golden/sc_z32_l12.pchk   terminated (3,6) SC-LDPC code, Z=32, L=12, w=3 (tools/gen_sc_pchk.py, seed 11)
  tests/golden/golden_sw.npz     per case: lratio, max_iter, win -> n, ok, packed dblk / pchk, sha256 of the final
                                 e->pr / e->lr arrays
The reference object keeps one .pchk per process, hence a script of its own."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_regular_pchk  # noqa: E402
import gen_sc_pchk  # noqa: E402
import oraclelib as ol  # noqa: E402

Z, L, SEED, W = 32, 12, 11, 3


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def cases(N):
    """(name, lratio, max_iter, win): all-zero codeword over a BSC / soft channel, incl. erasures and saturated ratios"""
    rs = np.random.RandomState(5)
    out = []
    for k, (eps, mi, win) in enumerate([(0.02, 20, 4), (0.05, 20, 4), (0.08, 30, 5), (0.11, 10, 3), (0.0, 5, 3), (0.06, 0, 4), (0.07, 50, 12)]):
        flips = rs.rand(N) < eps
        p = max(eps, 0.02)
        out.append(("bsc%d" % k, np.where(flips, p / (1 - p), (1 - p) / p), mi, win))
    llr = rs.normal(1.6, 1.3, N)
    llr[rs.rand(N) < 0.05] = 0.0
    out.append(("soft", np.exp(llr), 25, 4))
    sat = np.where(rs.rand(N) < 0.04, np.exp(-46.7), np.exp(46.7))
    out.append(("saturated", sat, 15, 4))
    return out


def main():
    gold = os.path.join(ROOT, "tests", "golden")
    path = os.path.join(gold, "sc_z32_l12.pchk")
    M, N, row_ptr, col_idx, Mv, Mc = gen_sc_pchk.gen_sc(Z, L, SEED)
    gen_regular_pchk.write_pchk(path, M, N, row_ptr, col_idx)
    ref = ol.RefLib(path)
    orc = ol.Oracle(path)
    out = {"names": [], "Mv": Mv, "Mc": Mc, "L": np.int32(L), "w": np.int32(W)}
    for name, lr, mi, win in cases(N):
        r = ref.decode_sw(lr, mi, L, W, win, Mv, Mc, want_msgs=True)
        o = orc.decode_sw(lr, mi, L, W, win, Mv, Mc, want_msgs=True)
        same = (r["n"] == o["n"] and r["ok"] == o["ok"] and np.array_equal(r["dblk"], o["dblk"]) and np.array_equal(r["pchk"], o["pchk"])
                and np.array_equal(r["pr"].view(np.uint64), o["pr"].view(np.uint64)) and np.array_equal(r["lr"].view(np.uint64), o["lr"].view(np.uint64)))
        print("  %-10s win=%2d max_iter=%2d  n=%2d ok=%d weight=%d  per-position %s  oracle==ref: %s"
              % (name, win, mi, r["n"], r["ok"], int(r["dblk"].sum()), o["iters_pos"].tolist(), same))
        out["names"].append(name)
        out[name + ".lratio"] = lr
        out[name + ".max_iter"] = np.int32(mi)
        out[name + ".win"] = np.int32(win)
        out[name + ".n"] = np.int32(r["n"])
        out[name + ".ok"] = np.int32(r["ok"])
        out[name + ".dblk"] = np.packbits(r["dblk"].astype(np.uint8), bitorder="little")
        out[name + ".pchk"] = np.packbits(r["pchk"].astype(np.uint8), bitorder="little")
        out[name + ".pr_sha"] = sha(r["pr"])
        out[name + ".lr_sha"] = sha(r["lr"])
    out["names"] = np.array(out["names"])
    np.savez_compressed(os.path.join(gold, "golden_sw.npz"), **out)
    print("wrote golden_sw.npz and", path)


if __name__ == "__main__":
    main()
