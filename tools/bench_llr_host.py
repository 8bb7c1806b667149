#!/usr/bin/env python3
"""fp64 LLR host batches through dnaldpc_decode_batch (147 KB per frame of the n=18432 code): frames/s and host->device
GB/s for several batch sizes, next to a device-resident run of the same frames (LR_F64 ratios in HBM) and to the input
copy alone. Answers "is this path bound by PCIe or by the decoder?".
  python tools/bench_llr_host.py [--frames 16384,65536] [--eps 0.006]"""
import argparse
import json
import os
import sys
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", default="16384,65536")
    ap.add_argument("--eps", type=float, default=0.006)
    ap.add_argument("--max-iter", type=int, default=100)
    a = ap.parse_args()
    import torch
    import _pkg
    import bench
    import oraclelib as ol
    ldpc = _pkg.load()
    code = ldpc.Code(ol.PCHK_18432)
    dec = ldpc.Decoder(code, devices=[0], wave_frames=4096)
    N, W = code.N, (code.N + 31) // 32
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream
    cws = ol.load_codewords()
    d_cw = torch.from_numpy(np.packbits(cws.astype(np.uint8), axis=1, bitorder="little").view(np.int32).copy()).to(dev)
    L = float(np.log((1 - a.eps) / a.eps))
    for F in [int(x) for x in a.frames.split(",")]:
        d_b = torch.empty((F, W), dtype=torch.int32, device=dev)
        dec.synth_bsc_device(d_cw.data_ptr(), 272, 1, 0, F, a.eps, d_b.data_ptr(), st)
        torch.cuda.synchronize()
        h_llr = torch.empty((F, N), dtype=torch.float64).pin_memory()
        for f0 in range(0, F, 4096):  # LLR = +-L from the received bits, built on the device in slices
            n = min(4096, F - f0)
            sh = torch.arange(32, device=dev, dtype=torch.int32)
            bits = ((d_b[f0:f0 + n].unsqueeze(-1) >> sh) & 1).reshape(n, W * 32)[:, :N]
            h_llr[f0:f0 + n].copy_(torch.where(bits == 0, L, -L).to(torch.float64))
        # the same frames resident in HBM as ratios
        d_lr = torch.exp(h_llr[:min(F, 32768)].to(dev))
        Fd = d_lr.shape[0]
        d_bits = torch.empty((Fd, W), dtype=torch.int32, device=dev)
        d_it = torch.empty(Fd, dtype=torch.int32, device=dev)
        d_ok = torch.empty(Fd, dtype=torch.uint8, device=dev)

        def go_dev():
            dec.decode_device(ldpc.IN_LR_F64, d_lr.data_ptr(), Fd, a.max_iter, bits_ptr=d_bits.data_ptr(), iters_ptr=d_it.data_ptr(),
                              ok_ptr=d_ok.data_ptr(), stream=st)
        go_dev()
        ms = bench._ev_time(torch, go_dev)
        del d_lr
        d_llr = h_llr[:Fd].to(dev)  # and as LLRs: exp() in the setup kernel

        def go_dev_llr():
            dec.decode_device(ldpc.IN_LLR_F64, d_llr.data_ptr(), Fd, a.max_iter, bits_ptr=d_bits.data_ptr(), iters_ptr=d_it.data_ptr(),
                              ok_ptr=d_ok.data_ptr(), stream=st)
        go_dev_llr()
        ms_llr = bench._ev_time(torch, go_dev_llr)
        del d_llr
        t_buf = torch.empty((min(F, 32768), N), dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        t_buf.copy_(h_llr[:t_buf.shape[0]], non_blocking=True)
        torch.cuda.synchronize()
        copy_gbs = t_buf.numel() * 8 / (time.perf_counter() - t0) / 1e9
        del t_buf
        r = bench.host_leg(ldpc, torch, dec, ldpc.IN_LLR_F64, h_llr, F, 0.0, a.max_iter, None)
        print(json.dumps({"frames": F, "host_frames_per_s": r["frames_per_s"], "host_h2d_gb_s": r["h2d_bytes"] / (r["ms"] * 1e-3) / 1e9,
                          "device_resident_frames": Fd, "device_resident_frames_per_s": Fd / (ms * 1e-3), "device_resident_llr_frames_per_s": Fd / (ms_llr * 1e-3),
                          "input_copy_alone_gb_s": copy_gbs, "avg_iters": r["avg_iters"]}))
        del h_llr, d_b


if __name__ == "__main__":
    main()
