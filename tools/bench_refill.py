#!/usr/bin/env python3
"""Converging-regime probe (frames that need a handful of iterations: every group admits frames in every tick).
Device-resident BSC batches of the n=18432 code; prints frames/s, frame-iterations/s, the whole-step fraction of the
measured HBM peak and the in-pipeline per-tick times (check pass / bit pass / scheduler) of ticks 8..71.
  python tools/bench_refill.py [--eps 0.006] [--frames 65536 1000000] [--c5]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--eps", type=float, default=0.006)
    ap.add_argument("--frames", type=int, nargs="+", default=[65536, 1000000])
    ap.add_argument("--max-iter", type=int, default=100)
    ap.add_argument("--wave", type=int, default=4096)
    ap.add_argument("--c5", action="store_true", help="also the n=65536 column-weight-3 code at eps 0.004")
    ap.add_argument("--tag", default="")
    ap.add_argument("--kind", default="bsc", choices=["bsc", "vote", "awgn"], help="input kind of the n=18432 probes (vote: configs[2], awgn: configs[3] at 4.6 dB)")
    a = ap.parse_args()
    import torch
    import _pkg
    ldpc = _pkg.load()
    peak, _ = bench.peaks()
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream
    res = {"tag": a.tag, "env": {k: v for k, v in os.environ.items() if k.startswith("DNALDPC_")}}

    def probe(dec, n, b_iter, F, eps, d_cw, n_cw, kind="bsc"):
        W = (n + 31) // 32
        if kind == "vote":
            d_in = torch.empty((F, n), dtype=torch.int8, device=dev)
            dec.synth_vote_device(d_cw.data_ptr(), n_cw, 7, 0, F, 3.9, 0.01, d_in.data_ptr(), st)
            k_in, param = ldpc.IN_VOTE_I8, 0.02
        elif kind == "awgn":
            param = ldpc.std_dev(4.6, 1 - 2048 / 18432)
            d_in = torch.empty((F, n), dtype=torch.float32, device=dev)
            dec.synth_awgn_device(d_cw.data_ptr(), n_cw, 7, 0, F, param, d_in.data_ptr(), st)
            k_in = ldpc.IN_AWGN_F32
        else:
            d_in = torch.empty((F, W), dtype=torch.int32, device=dev)
            dec.synth_bsc_device(d_cw.data_ptr() if d_cw is not None else None, n_cw, 7, 0, F, eps, d_in.data_ptr(), st)
            k_in, param = ldpc.IN_BSC_BITS, eps
        d_bits = torch.empty((F, W), dtype=torch.int32, device=dev)
        d_it = torch.empty(F, dtype=torch.int32, device=dev)
        d_ok = torch.empty(F, dtype=torch.uint8, device=dev)

        def go():
            dec.decode_device(k_in, d_in.data_ptr(), F, a.max_iter, param=param, bits_ptr=d_bits.data_ptr(),
                              iters_ptr=d_it.data_ptr(), ok_ptr=d_ok.data_ptr(), stream=st)
        go()
        dec.set_profiling(2)
        go()
        torch.cuda.synchronize()
        tr = dec.trace()
        dec.set_profiling(0)
        ms = min(bench._ev_time(torch, go) for _ in range(2))
        fi = float(d_it.sum().item())
        return {"frames": F, "eps": eps, "ms": round(ms, 2), "frames_per_s": round(F / ms * 1e3), "frame_iters_per_s": round(fi / ms * 1e3),
                "avg_iters": round(fi / F, 3), "frac": round(b_iter * fi / (ms * 1e-3) / 1e9 / peak, 4),
                "tick_ms": {"row": round(tr[0], 3), "col": round(tr[1], 3), "sched": round(tr[2], 3), "ticks": tr[3]},
                "launches": dec.stats()["kernel_launches"], "compactions": dec.stats()["compactions"]}

    code = ldpc.Code(bench.PCHK)
    dec = ldpc.Decoder(code, devices=[0], wave_frames=a.wave)
    cw = np.fromfile(bench.CW_BITS, dtype=np.uint8).view(np.int32).reshape(272, -1)
    d_cw = torch.from_numpy(cw).to(dev)
    for F in a.frames:
        res["n18432_%s_%d" % (a.kind, F)] = probe(dec, bench.N, bench.B_ITER, F, a.eps, d_cw, 272, a.kind)
    dec.close()
    if a.c5:
        import gen_regular_pchk
        n5, m5 = 65536, 6554
        row_ptr, col_idx = gen_regular_pchk.gen_regular(n5, m5, 3, 5)
        code5 = ldpc.Code(csr=(m5, n5, row_ptr, col_idx))
        dec5 = ldpc.Decoder(code5, devices=[0], wave_frames=a.wave)
        res["c5_n65536_32768"] = probe(dec5, n5, 32 * code5.E + 8.25 * n5, 32768, 0.004, None, 0)
        res["c5_n65536_steady_8192"] = probe(dec5, n5, 32 * code5.E + 8.25 * n5, 8192, 0.02, None, 0)
        dec5.close()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
