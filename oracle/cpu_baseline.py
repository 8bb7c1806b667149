#!/usr/bin/env python3
"""CPU baseline timing of the reference's BP decode path (TEST/BENCH INFRASTRUCTURE, never the product).

Times Run_Belief_Propagation_Decoder (dec.cpp:583-605) from the UNMODIFIED reference build oracle/_ref/libldpc_ref.so
("kind": "reference"), or the C restatement oracle/liboracle.so ("port") where the reference build is absent,
one process per host core, each pinned, each decoding its own slice of the same synthetic workload bench.py uses
(frame f = codeword[f % 272] through a BSC(eps) drawn from the shared counter RNG). Prints one JSON object.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))


def _worker(args):
    rank, core, pchk, frames, eps, max_iter, seed, use_ref = args
    try:
        os.sched_setaffinity(0, {core})
    except Exception:
        pass
    import oraclelib as ol
    dec = ol.RefLib(pchk) if use_ref else ol.Oracle(pchk)
    N = dec.N
    if N == 18432:
        cws = ol.load_codewords()
    else:
        cws = np.zeros((1, N), np.int8)
    lr0, lr1 = (1 - eps) / eps, eps / (1 - eps)
    lr = np.zeros((len(frames), N))
    for k, f in enumerate(frames):
        recv = cws[f % len(cws)] ^ ol.bsc_flips(seed, f, N, eps)
        lr[k] = np.where(recv == 0, lr0, lr1)
    t0 = time.perf_counter()
    r = dec.decode_many(lr, max_iter)
    dt = time.perf_counter() - t0
    return dt, int(r["n"].sum()), len(frames)


def run(pchk, procs, frames_per_proc, eps, max_iter, seed=7, frame0=0):
    import oraclelib as ol
    use_ref = ol.RefLib.available()
    if not use_ref:
        ol.build_oracle()
    try:
        cores = sorted(os.sched_getaffinity(0))
    except Exception:
        cores = list(range(os.cpu_count() or 1))
    procs = min(procs, len(cores)) if procs > 0 else len(cores)
    jobs = []
    for r in range(procs):
        fr = list(range(frame0 + r * frames_per_proc, frame0 + (r + 1) * frames_per_proc))
        jobs.append((r, cores[r % len(cores)], pchk, fr, eps, max_iter, seed, use_ref))
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        # warm the pool (imports, .pchk parse) outside the timed region; each worker times only its decode loop
        res = pool.map(_worker, jobs)
    wall = time.perf_counter() - t0
    tmax = max(r[0] for r in res)
    frames = sum(r[2] for r in res)
    iters = sum(r[1] for r in res)
    N = 18432 if "18432" in os.path.basename(pchk) else None
    return dict(kind="reference" if use_ref else "port", cores=procs, frames=frames, frame_iters=iters,
                decode_s=tmax, wall_s=wall, ms_per_iter_per_core=1e3 * sum(r[0] for r in res) / max(iters, 1),
                frames_per_s=frames / tmax, eps=eps, max_iter=max_iter, N=N)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--pchk", default=os.path.join(os.path.dirname(HERE), "tests", "golden", "decode_n18432_m2048_final.pchk"))
    ap.add_argument("--procs", type=int, default=0)
    ap.add_argument("--frames-per-proc", type=int, default=2)
    ap.add_argument("--eps", type=float, default=0.02)
    ap.add_argument("--max-iter", type=int, default=100)
    a = ap.parse_args()
    print(json.dumps(run(a.pchk, a.procs, a.frames_per_proc, a.eps, a.max_iter)))
