#!/usr/bin/env python3
"""How fast does a pinned host->device copy run while the flooding decoder saturates HBM? (copy engine vs SM traffic)"""
import os, sys, threading, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, _pkg, oraclelib as ol
ldpc = _pkg.load()
code = ldpc.Code(ol.PCHK_18432)
dec = ldpc.Decoder(code, devices=[0], wave_frames=4096)
dev = torch.device("cuda", 0)
N, W, F = code.N, (code.N + 31) // 32, 16384
st = torch.cuda.current_stream().cuda_stream
d_in = torch.empty((F, W), dtype=torch.int32, device=dev)
dec.synth_bsc_device(None, 0, 1, 0, F, 0.02, d_in.data_ptr(), st)
d_bits = torch.empty((F, W), dtype=torch.int32, device=dev); d_it = torch.empty(F, dtype=torch.int32, device=dev); d_ok = torch.empty(F, dtype=torch.uint8, device=dev)
def decode():
    dec.decode_device(ldpc.IN_BSC_BITS, d_in.data_ptr(), F, 100, param=0.02, bits_ptr=d_bits.data_ptr(), iters_ptr=d_it.data_ptr(), ok_ptr=d_ok.data_ptr(), stream=st)
decode(); torch.cuda.synchronize()
h = torch.empty(4 << 30, dtype=torch.uint8).pin_memory()
d = torch.empty(4 << 30, dtype=torch.uint8, device=dev)
cs = torch.cuda.Stream()
def copy_gbs(chunk):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(cs):
        for o in range(0, h.numel(), chunk):
            d[o:o + chunk].copy_(h[o:o + chunk], non_blocking=True)
    cs.synchronize()
    return h.numel() / (time.perf_counter() - t0) / 1e9
idle = {c: copy_gbs(c) for c in (32 << 20, 256 << 20)}
th = threading.Thread(target=decode); th.start()
time.sleep(0.15)
load = {c: copy_gbs(c) for c in (32 << 20, 256 << 20)}
alive = th.is_alive()
th.join()
print(json.dumps({"copy_gb_s_idle": idle, "copy_gb_s_while_decoding": load, "decoder_still_running_after_copies": alive}))
