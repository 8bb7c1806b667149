#!/usr/bin/env python3
"""Re-decoding sweep (ex_decoder/decoder.py:594-664) on vote-count inputs: dnaldpc_redecode_sweep_ex, whose inputs stay in
HBM after round 0 (later rounds run over device-side row lists; only that round's results cross PCIe), against the same
rounds driven from the host the way round 1 of this repo did it (failed frames gathered on the host and uploaded again
as a fresh batch per round). Prints times, the bytes each path moves host->device and checks that both agree.
  python tools/bench_sweep.py [--frames 65536]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _pkg  # noqa: E402
import oraclelib as ol  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=65536)
    ap.add_argument("--max-iter", type=int, default=25)
    a = ap.parse_args()
    import torch
    ldpc = _pkg.load()
    code = ldpc.Code(ol.PCHK_18432)
    dec = ldpc.Decoder(code, devices=[0], wave_frames=4096)
    N, F = 18432, a.frames
    cws = ol.load_codewords()
    cw_packed = np.packbits(cws.astype(np.uint8), axis=1, bitorder="little").view(np.uint32)
    d_cw = torch.from_numpy(cw_packed.astype(np.int32)).cuda()
    d_k = torch.empty((F, N), dtype=torch.int8, device="cuda")
    dec.synth_vote_device(d_cw.data_ptr(), 272, 12, 0, F, 3.9, 0.03, d_k.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    k = d_k.cpu().numpy()
    del d_k
    eps_rounds = [0.20, 0.10, 0.07, 0.05]
    dec.redecode_sweep_ex(ldpc.IN_VOTE_I8, k[:4096], a.max_iter, eps_rounds)  # warm-up
    t0 = time.perf_counter()
    r = dec.redecode_sweep_ex(ldpc.IN_VOTE_I8, k, a.max_iter, eps_rounds)
    t_res = time.perf_counter() - t0
    # host-driven rounds
    t0 = time.perf_counter()
    ok = np.zeros(F, bool)
    rounds = np.zeros(F, np.int32)
    bits = np.zeros((F, N), np.int8)
    h2d = 0
    for rd, e in enumerate(eps_rounds):
        todo = np.nonzero(~ok)[0]
        if len(todo) == 0:
            break
        sub = np.ascontiguousarray(k[todo])
        h2d += sub.nbytes
        one = dec.decode(ldpc.IN_VOTE_I8, sub, a.max_iter, param=e)
        bits[todo] = one["bits"]
        ok[todo] = one["ok"] == 1
        rounds[todo] = rd
    t_host = time.perf_counter() - t0
    per_round = [int((r["rounds"] >= rd).sum()) for rd in range(len(eps_rounds))]
    out = {"frames": F, "eps_rounds": eps_rounds, "max_iter": a.max_iter, "frames_per_round": per_round,
           "resident_sweep_s": t_res, "host_driven_rounds_s": t_host, "speedup": t_host / t_res,
           "h2d_bytes_resident": int(k.nbytes), "h2d_bytes_host_driven": int(h2d),
           "identical": bool(np.array_equal(r["rounds"], rounds) and np.array_equal(r["bits"], bits) and np.array_equal(r["ok"] == 1, ok))}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
