#!/usr/bin/env python3
"""Throughput of the sliding-window decoder (dnaldpc_decode_window = Run_SW_Decoder, dec.cpp:2092-2196) on a terminated
(3,6) spatially-coupled code from tools/gen_sc_pchk.py (Z x L), BSC inputs, host buffers. Prints frames/s, window
updates per frame and the algorithmic GB/s of the update kernels: per update of a position the check kernel reads pr and
writes lr for every edge of the window's checks (16 B per edge) and the bit kernel reads lr and rewrites pr for every
in-window edge of the window's bits (24 B per edge) plus the channel ratio (8 B per bit).
  python tools/bench_sw.py [--z 1024] [--L 24] [--win 6] [--frames 8192] [--eps 0.03] [--max-iter 20]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import _pkg  # noqa: E402
import gen_sc_pchk  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--z", type=int, default=1024)
    ap.add_argument("--L", type=int, default=24)
    ap.add_argument("--win", type=int, default=6)
    ap.add_argument("--frames", type=int, default=8192)
    ap.add_argument("--eps", type=float, default=0.03)
    ap.add_argument("--max-iter", type=int, default=20)
    ap.add_argument("--wave", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--warm", type=int, default=0, help="frames of the warm-up call (default: one wave; negative: no warm-up call, e.g. under ncu)")
    a = ap.parse_args()
    ldpc = _pkg.load()
    M, N, row_ptr, col_idx, Mv, Mc = gen_sc_pchk.gen_sc(a.z, a.L, 11)
    code = ldpc.Code(csr=(M, N, row_ptr, col_idx))
    dec = ldpc.Decoder(code, devices=[0], wave_frames=a.wave)
    rs = np.random.RandomState(1)
    flips = rs.rand(a.frames, N) < a.eps          # all-zero codeword through a BSC
    lr = np.where(flips, a.eps / (1 - a.eps), (1 - a.eps) / a.eps)
    D = a.L + 3 - 1
    import torch
    C = ldpc.C
    lr = torch.from_numpy(lr).pin_memory()       # pinned host buffers: the copies run at PCIe speed
    W = (N + 31) // 32
    bits = torch.zeros((a.frames, W), dtype=torch.int32).pin_memory()
    iters = torch.zeros(a.frames, dtype=torch.int32).pin_memory()
    okf = torch.zeros(a.frames, dtype=torch.uint8).pin_memory()
    mv = np.ascontiguousarray(Mv[:D], dtype=np.int32); mc = np.ascontiguousarray(Mc[:D], dtype=np.int32)
    wd = ldpc.Window(code_type=0, L=a.L, w=3, win=a.win, Mv=mv.ctypes.data, Mc=mc.ctypes.data)
    out = ldpc.Output(bits=bits.data_ptr(), dblk=None, iters=iters.data_ptr(), is_codeword=okf.data_ptr(), posterior=None, pchk=None)

    def call(nf):  # the C-ABI call alone (host buffers in, host buffers out); no numpy post-processing in the timed region
        rc = ldpc.lib().dnaldpc_decode_window(dec._h, C.byref(wd), lr.data_ptr(), nf, a.max_iter, C.byref(out))
        if rc:
            raise RuntimeError(ldpc.lib().dnaldpc_last_error())
    if a.warm >= 0:
        call(min(a.frames, a.warm or a.wave))  # warm-up: slot arrays, staging buffers
    dts = []
    for _ in range(a.reps):
        t0 = time.perf_counter()
        call(a.frames)
        dts.append(time.perf_counter() - t0)
    dt = min(dts)
    # the input copy alone (pinned host -> device), for comparison: the call cannot be faster than this
    dev_buf = torch.empty_like(lr, device="cuda")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dev_buf.copy_(lr, non_blocking=True)
    torch.cuda.synchronize()
    h2d_s = time.perf_counter() - t0
    del dev_buf
    st = dec.stats()
    r = {"iters": iters.numpy(), "ok": okf.numpy(),
         "bits": np.unpackbits(bits.numpy().view(np.uint8), axis=1, bitorder="little")[:, :N]}
    # updates per frame = L positions x (n + 1); iters = floor(sum n / L)
    upd = float((r["iters"].astype(np.float64) + 1).sum()) * a.L
    edges_win = a.win * 2 * a.z * 3               # edges of the bits of one full window
    bytes_upd = 16.0 * (a.win + 2) * a.z * 6 + 24.0 * edges_win + 8.0 * a.win * 2 * a.z
    out = {"code": {"Z": a.z, "L": a.L, "N": N, "M": M, "E": int(len(col_idx)), "win": a.win}, "frames": a.frames, "eps": a.eps,
           "max_iter": a.max_iter, "seconds": dt, "frames_per_s": a.frames / dt, "decoded_gbit_s": a.frames * N / dt / 1e9,
           "fer": 1.0 - float((r["ok"] == 1).mean()), "bit_errors": int(r["bits"].sum()),
           "avg_updates_per_position": upd / a.frames / a.L, "kernel_launches": st["kernel_launches"], "ticks": st["waves"], "wave_frames": a.wave,
           "schedule": "lock-step" if os.environ.get("DNALDPC_SW_LOCKSTEP", "0") not in ("", "0") else "groups",
           "algorithmic_gb_s": upd * bytes_upd / dt / 1e9,
           "input_bytes": int(lr.numel()) * 8, "input_copy_alone_s": h2d_s, "input_copy_alone_gb_s": int(lr.numel()) * 8 / h2d_s / 1e9}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
