#!/usr/bin/env python3
"""Secondary throughput probe for the BASELINE configs that are NOT the bench line (bench.py measures configs[1] only):
any code (the n=18432 matrix or the config-5 random regular code n=65536 / column weight 3 / rate 0.9), BSC inputs
generated on the device, device-resident buffers, CUDA-event timing. Prints one JSON object.

  python tools/bench_code.py --code n65536 --frames 16384 --eps 0.003 --max-iter 60
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--code", default="n65536", choices=["n18432", "n65536"])
    ap.add_argument("--frames", type=int, default=16384)
    ap.add_argument("--eps", type=float, default=0.003)
    ap.add_argument("--max-iter", type=int, default=60)
    ap.add_argument("--wave", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--alg", default="bp", choices=["bp", "minsum"])
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
    a = ap.parse_args()

    import torch
    import _pkg
    ldpc = _pkg.load()
    if a.code == "n18432":
        code = ldpc.Code(os.path.join(ROOT, "tests", "golden", "decode_n18432_m2048_final.pchk"))
    else:
        import gen_regular_pchk
        row_ptr, col_idx = gen_regular_pchk.gen_regular(65536, 6554, 3, 5)
        code = ldpc.Code(csr=(6554, 65536, row_ptr, col_idx))
    N, M, E = code.N, code.M, code.E
    dec = ldpc.Decoder(code, devices=[0], wave_frames=a.wave, precision=ldpc.PREC_F32 if a.precision == "f32" else ldpc.PREC_F64)
    W = (N + 31) // 32
    F = a.frames
    dev = torch.device("cuda", 0)
    d_in = torch.empty((F, W), dtype=torch.int32, device=dev)
    d_bits = torch.empty((F, W), dtype=torch.int32, device=dev)
    d_it = torch.empty(F, dtype=torch.int32, device=dev)
    d_ok = torch.empty(F, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    dec.synth_bsc_device(None, 0, 11, 0, F, a.eps, d_in.data_ptr(), st)   # all-zero codeword through a BSC
    flags = ldpc.FLAG_MINSUM if a.alg == "minsum" else 0
    C = ldpc.C
    inp = ldpc.Input(kind=ldpc.IN_BSC_BITS, flags=flags, data=d_in.data_ptr(), frame_stride=0, param=a.eps, table=None)
    out = ldpc.Output(bits=d_bits.data_ptr(), dblk=None, iters=d_it.data_ptr(), is_codeword=d_ok.data_ptr(), posterior=None, pchk=None)

    def step():
        rc = ldpc.lib().dnaldpc_decode_batch_device(dec._h, C.byref(inp), F, a.max_iter, C.byref(out), st)
        if rc:
            raise RuntimeError(ldpc.lib().dnaldpc_last_error())
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    fi = int(d_it.sum().item())
    esz = 4 if a.precision == "f32" else 8
    b_iter = (4 * E + N) * esz + N / 4          # 32E + 8.25N for fp64
    ok = d_ok.cpu().numpy()
    good_zero = bool((d_bits[torch.from_numpy(ok.astype(bool)).to(dev)] == 0).all()) if ok.any() else True
    print(json.dumps({
        "code": a.code, "N": N, "M": M, "E": E, "frames": F, "eps": a.eps, "max_iter": a.max_iter, "algorithm": a.alg,
        "dtype": a.precision, "ms_per_step": ms, "decoded_gbit_s": F * N / (ms * 1e-3) / 1e9,
        "frame_iters_per_s": fi / (ms * 1e-3), "avg_iters": fi / F, "fer": float(1 - ok.mean()),
        "algorithmic_gb_s": b_iter * fi / (ms * 1e-3) / 1e9, "b_iter_bytes": b_iter,
        "converged_frames_are_the_sent_codeword": good_zero}))


if __name__ == "__main__":
    main()
