#!/usr/bin/env python3
"""Golden digests for BASELINE configs[0] from the UNMODIFIED reference (run in the build container only):
all 272 ex_decoder codewords, frame f = codeword[f] through the synthetic BSC(eps) of the shared counter RNG (seed 7,
frame index f), prprp max 100 iterations, decoded by Run_Belief_Propagation_Decoder (dec.cpp:583-605) of
oracle/_ref/libldpc_ref.so, one process per host core. eps = 0.02 is the BASELINE operating point (no frame converges,
SURVEY 6.2); eps = 0.0075 is the same run in the waterfall, where the iteration counts differ frame by frame.
Writes tests/golden/golden_c0.npz: per eps and frame the iteration count n, the success flag, sha256(dblk bytes) and
sha256(posterior float64 bytes). The GPU box never runs this script.
"""
import hashlib
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
SEED, MAX_ITER, EPS = 7, 100, (0.02, 0.0075)


def work(args):
    eps, frames = args
    import oraclelib as ol
    ref = ol.RefLib(ol.PCHK_18432)
    cws = ol.load_codewords()
    out = []
    for f in frames:
        recv = cws[f] ^ ol.bsc_flips(SEED, f, ref.N, eps)
        lr = np.where(recv == 0, (1 - eps) / eps, eps / (1 - eps))
        r = ref.decode(lr, MAX_ITER, want_post=True)
        out.append((f, int(r["n"]), int(r["ok"]), hashlib.sha256(r["dblk"].astype(np.uint8).tobytes()).digest(),
                    hashlib.sha256(r["post"].tobytes()).digest()))
    return out


def main():
    import oraclelib as ol
    assert ol.RefLib.available(), "build oracle/_ref first (make -C oracle ref)"
    nproc = len(os.sched_getaffinity(0))
    res = {}
    with mp.get_context("spawn").Pool(nproc) as pool:
        for eps in EPS:
            jobs = [(eps, list(range(r, 272, nproc))) for r in range(nproc)]
            rows = sorted(sum(pool.map(work, jobs), []))
            tag = "eps%g" % eps
            res[tag + ".n"] = np.array([r[1] for r in rows], np.int32)
            res[tag + ".ok"] = np.array([r[2] for r in rows], np.uint8)
            res[tag + ".dblk_sha"] = np.frombuffer(b"".join(r[3] for r in rows), np.uint8).reshape(272, 32)
            res[tag + ".post_sha"] = np.frombuffer(b"".join(r[4] for r in rows), np.uint8).reshape(272, 32)
            print(tag, "iterations:", np.bincount(res[tag + ".n"]).nonzero()[0].tolist(), "converged:", int(res[tag + ".ok"].sum()))
    res["seed"] = np.int64(SEED)
    res["max_iter"] = np.int32(MAX_ITER)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_c0.npz"), **res)


if __name__ == "__main__":
    main()
