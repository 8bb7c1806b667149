#!/usr/bin/env python3
"""Randomised campaign of the GPU decoder (through the C ABI) against the CPU oracle on the small N=120 code, where the
oracle is cheap: random batch sizes, slot counts, iteration caps, input classes, sum-product and min-sum; every frame's
n, success flag, decoded bits, syndrome and posterior bit for bit. Exercises the slot scheduler (admission, two-round
refill, drain-tail compaction, kernel-variant choices) far beyond the test suite. Needs a GPU.

  fuzz_gpu_vs_oracle.py SECONDS        flooding sum-product + min-sum, small code
  fuzz_gpu_vs_oracle.py SECONDS sw     sliding-window BP on the SC-LDPC code (window 1..14, code_type 0 / 1)
  fuzz_gpu_vs_oracle.py SECONDS big    the n=18432 code (tensor-memory check pass, compact gathers of starting lanes): random
                                       batch sizes / slot counts / iteration caps / eps mixes, host and device buffers,
                                       a random sample of every batch (stragglers included) against the oracle
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import _pkg  # noqa: E402
import oraclelib as ol  # noqa: E402
from fuzz_oracle_vs_ref import ratios  # noqa: E402


def main_sw(seconds):
    import gen_sc_pchk
    ldpc = _pkg.load()
    path = os.path.join(ol.GOLDEN, "sc_z32_l12.pchk")
    _, N, _, _, Mv, Mc = gen_sc_pchk.gen_sc(32, 12, 11)
    code, orc = ldpc.Code(path), ol.Oracle(path)
    rs = np.random.RandomState(5)
    decs = {}
    t0, frames, batches = time.time(), 0, 0
    while time.time() - t0 < seconds:
        wave = int(rs.choice([32, 64, 256]))
        F = int(rs.choice([1, 33, 100, 300]))
        ct = int(rs.randint(0, 2))
        D = 12 + 3 - 1 if ct == 0 else 12 + (3 - 1) // 2
        win, mi = min(int(rs.randint(1, 15)), D), int(rs.choice([0, 1, 2, 5, 12, 40]))
        dec = decs.get(wave) or decs.setdefault(wave, ldpc.Decoder(code, wave_frames=wave))
        q = rs.choice([0.0, 0.02, 0.04, 0.06, 0.08, 0.1, 0.15], size=F)
        kinds = rs.randint(0, 3, size=F)
        lr = np.stack([ratios(rs, N, q[f], kinds[f]) for f in range(F)])
        r = dec.decode_window(lr, mi, 12, 3, win, Mv[:D], Mc[:D], code_type=ct, want=("bits", "iters", "ok", "pchk"))
        for f in range(F):
            o = orc.decode_sw(lr[f], mi, 12, 3, win, Mv[:D], Mc[:D], code_type=ct)
            assert r["iters"][f] == o["n"] and r["ok"][f] == o["ok"], (wave, F, win, mi, ct, f)
            assert np.array_equal(r["bits"][f], o["dblk"]) and np.array_equal(r["pchk"][f].astype(np.int8), o["pchk"]), (wave, F, win, mi, ct, f)
        frames += F
        batches += 1
    print("GPU sliding window vs oracle: %d batches, %d frames identical" % (batches, frames))


def main_big(seconds):
    import torch
    ldpc = _pkg.load()
    code, orc = ldpc.Code(ol.PCHK_18432), ol.Oracle(ol.PCHK_18432)
    cws = ol.load_codewords()
    N, W = 18432, 576
    rs = np.random.RandomState(2024)
    decs = {}
    t0, frames, checked, batches = time.time(), 0, 0, 0
    while time.time() - t0 < seconds:
        wave = int(rs.choice([64, 256, 1024, 4096]))
        F = int(rs.choice([40, 333, 1500, 6000]))
        mi = int(rs.choice([3, 8, 20, 50]))
        mix = [(0.004, 0.006), (0.006, 0.0075, 0.02), (0.0075,), (0.003, 0.02)][int(rs.randint(0, 4))]
        seed = int(rs.randint(1, 1 << 30))
        dec = decs.get(wave) or decs.setdefault(wave, ldpc.Decoder(code, wave_frames=wave))
        eps = np.array([mix[f % len(mix)] for f in range(F)])
        recv = np.stack([cws[f % 272] ^ ol.bsc_flips(seed, f, N, eps[f]) for f in range(F)])
        lr = np.where(recv == 0, ((1 - eps) / eps)[:, None], (eps / (1 - eps))[:, None])
        if rs.randint(0, 2):
            r = dec.decode(ldpc.IN_LR_F64, lr, mi, want=("bits", "iters", "ok", "post"))
            it, ok, bits, post = r["iters"], r["ok"], r["bits"], r["post"]
        else:
            d_lr = torch.from_numpy(lr).cuda()
            d_bits = torch.zeros((F, W), dtype=torch.int32, device="cuda")
            d_it = torch.zeros(F, dtype=torch.int32, device="cuda")
            d_ok = torch.zeros(F, dtype=torch.uint8, device="cuda")
            d_post = torch.zeros((F, N), dtype=torch.float64, device="cuda")
            dec.decode_device(ldpc.IN_LR_F64, d_lr.data_ptr(), F, mi, bits_ptr=d_bits.data_ptr(), iters_ptr=d_it.data_ptr(),
                              ok_ptr=d_ok.data_ptr(), post_ptr=d_post.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            it, ok, post = d_it.cpu().numpy(), d_ok.cpu().numpy(), d_post.cpu().numpy()
            bits = np.unpackbits(d_bits.cpu().numpy().view(np.uint8).reshape(F, W * 4), axis=1, bitorder="little")[:, :N].astype(np.int8)
        assert dec.stats()["frames"] == F and dec.stats()["frame_iters"] == int(it.sum())
        sample = set(rs.choice(F, size=min(F, 6), replace=False).tolist())
        sample |= set(np.nonzero(it == it.max())[0][:2].tolist()) | set(np.nonzero(it == it.min())[0][:1].tolist())
        for f in sorted(sample):
            o = orc.decode(lr[f], mi)
            assert it[f] == o["n"] and ok[f] == o["ok"] and np.array_equal(bits[f], o["dblk"]), (wave, F, mi, mix, seed, f)
            assert np.array_equal(post[f].view(np.uint64), o["post"].view(np.uint64)), ("post", wave, F, mi, mix, seed, f)
        good = ok == 1
        assert np.array_equal(bits[good], cws[np.arange(F)[good] % 272])  # every frame flagged as a codeword is the sent one
        frames += F
        checked += len(sample)
        batches += 1
    print("GPU n=18432 vs oracle: %d batches, %d frames (all flagged codewords = the sent ones), %d frames bit for bit incl. posteriors"
          % (batches, frames, checked))


def main():
    seconds = float(sys.argv[1])
    if len(sys.argv) > 2 and sys.argv[2] == "sw":
        return main_sw(seconds)
    if len(sys.argv) > 2 and sys.argv[2] == "big":
        return main_big(seconds)
    ldpc = _pkg.load()
    path = os.path.join(ol.GOLDEN, "small_n120_m60.pchk")
    code, orc = ldpc.Code(path), ol.Oracle(path)
    N = code.N
    rs = np.random.RandomState(99)
    t0, frames, batches, comp = time.time(), 0, 0, 0
    decs = {}
    while time.time() - t0 < seconds:
        wave = int(rs.choice([32, 64, 96, 256, 1024, 4096]))
        F = int(rs.choice([1, 31, 33, 100, 700, 3000, 9000]))
        mi = int(rs.choice([0, 1, 4, 20, 60, 200]))
        dec = decs.get(wave) or decs.setdefault(wave, ldpc.Decoder(code, wave_frames=wave))
        q = rs.choice([0.0, 0.02, 0.05, 0.08, 0.12, 0.2], size=F)
        kinds = rs.randint(0, 4, size=F)
        lr = np.stack([ratios(rs, N, q[f], kinds[f]) for f in range(F)])
        r = dec.decode(ldpc.IN_LR_F64, lr, mi, want=("bits", "iters", "ok", "post", "pchk"))
        comp += dec.stats()["compactions"]
        llr = np.log(np.maximum(lr, 1e-300))
        llr = np.where(np.isfinite(llr), llr, 700.0)
        m = dec.decode(ldpc.IN_LLR_F64, llr, mi, flags=ldpc.FLAG_MINSUM, want=("bits", "iters", "ok", "post"))
        comp += dec.stats()["compactions"]
        for f in range(F):
            o = orc.decode(lr[f], mi)
            assert r["iters"][f] == o["n"] and r["ok"][f] == o["ok"], ("bp", wave, F, mi, f)
            assert np.array_equal(r["bits"][f], o["dblk"]) and np.array_equal(r["pchk"][f].astype(np.int8), o["pchk"]), ("bp", wave, F, mi, f)
            assert np.array_equal(r["post"][f].view(np.uint64), o["post"].view(np.uint64)), ("bp post", wave, F, mi, f)
            o = orc.decode_minsum(llr[f], mi)
            assert m["iters"][f] == o["n"] and m["ok"][f] == o["ok"] and np.array_equal(m["bits"][f], o["dblk"]), ("ms", wave, F, mi, f)
            assert np.array_equal(m["post"][f].view(np.uint64), o["post"].view(np.uint64)), ("ms post", wave, F, mi, f)
        frames += F
        batches += 1
    print("GPU vs oracle: %d batches, %d frames x 2 algorithms identical; %d compactions" % (batches, frames, comp))


if __name__ == "__main__":
    main()
