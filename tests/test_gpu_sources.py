"""How batches reach the engines (csrc/sources.cu), on the GPU through the C ABI: host batches staged through rings while
the engine decodes, device-resident batches, re-decoding sweeps over resident inputs, the device input generators, and
BASELINE configs[0] / [2] / [3] at their stated sizes."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

import _pkg
import gen_regular_pchk
import oraclelib as ol

pytestmark = pytest.mark.gpu
ldpc = _pkg.load()
N, M, W = 18432, 2048, 576


@pytest.fixture(scope="module")
def code():
    return ldpc.Code(ol.PCHK_18432)


@pytest.fixture(scope="module")
def orc():
    return ol.Oracle(ol.PCHK_18432)


@pytest.fixture(scope="module")
def cws():
    return ol.load_codewords()


def _mixed_frames(cws, F, seed, eps_list=(0.004, 0.006, 0.0075, 0.02)):
    eps = np.array([eps_list[f % len(eps_list)] for f in range(F)])
    recv = np.stack([cws[f % 272] ^ ol.bsc_flips(seed, f, N, eps[f]) for f in range(F)])
    lr = np.where(recv == 0, ((1 - eps) / eps)[:, None], (eps / (1 - eps))[:, None])
    return recv, lr, eps


def test_golden_config0_all_272_frames(code):
    """BASELINE configs[0]: all 272 ex_decoder codewords + synthetic BSC, prprp max 100 iterations, against digests of
    the UNMODIFIED reference's outputs (tests/golden/golden_c0.npz, generator tools/make_golden_c0.py): iteration counts,
    success flags, decoded bits and posteriors (sha256 of the raw float64 bytes) of every frame, at the BASELINE point
    eps = 0.02 and in the waterfall (eps = 0.0075)."""
    g = np.load(os.path.join(ol.GOLDEN, "golden_c0.npz"))
    cws = ol.load_codewords()
    dec = ldpc.Decoder(code, wave_frames=512)
    for eps in (0.02, 0.0075):
        tag = "eps%g" % eps
        recv = np.stack([cws[f] ^ ol.bsc_flips(int(g["seed"]), f, N, eps) for f in range(272)])
        lr = np.where(recv == 0, (1 - eps) / eps, eps / (1 - eps))
        r = dec.decode(ldpc.IN_LR_F64, lr, int(g["max_iter"]), want=("dblk", "iters", "ok", "post"))
        assert np.array_equal(r["iters"], g[tag + ".n"]), eps
        assert np.array_equal(r["ok"], g[tag + ".ok"]), eps
        for f in range(272):
            assert hashlib.sha256(r["dblk"][f].tobytes()).digest() == g[tag + ".dblk_sha"][f].tobytes(), (eps, f)
            assert hashlib.sha256(r["post"][f].tobytes()).digest() == g[tag + ".post_sha"][f].tobytes(), (eps, f)
        # the packed-bit BSC input kind sees the same frames
        packed = np.packbits(recv.astype(np.uint8), axis=1, bitorder="little").view(np.uint32)
        b = dec.decode(ldpc.IN_BSC_BITS, packed, int(g["max_iter"]), param=eps)
        assert np.array_equal(b["iters"], r["iters"]) and np.array_equal(b["bits"], r["dblk"].astype(np.int8))
    dec.close()


def test_host_batch_streams_through_rings(code, orc, cws):
    """A host batch much larger than the slots and the input ring (3000 frames, 256 slots): chunks are copied in, decoded
    and copied back concurrently. Every output kind, pageable and pinned memory, a strided input; identical to the
    device-resident path and to the oracle on a sample."""
    import torch
    F = 3000
    recv, lr, eps = _mixed_frames(cws, F, 91)
    dec = ldpc.Decoder(code, wave_frames=256)
    want = ("bits", "dblk", "iters", "ok", "post", "pchk")
    a = dec.decode(ldpc.IN_LR_F64, lr, 30, want=want)
    for f in list(range(0, F, 397)) + [F - 1]:
        o = orc.decode(lr[f], 30)
        assert a["iters"][f] == o["n"] and a["ok"][f] == o["ok"], f
        assert np.array_equal(a["bits"][f], o["dblk"]) and np.array_equal(a["dblk"][f], o["dblk"].astype(np.uint8)), f
        assert np.array_equal(a["post"][f].view(np.uint64), o["post"].view(np.uint64)), f
        assert np.array_equal(a["pchk"][f], o["pchk"].astype(np.uint8)), f
    assert dec.stats()["frames"] == F and dec.stats()["frame_iters"] == int(a["iters"].sum())
    # device-resident inputs and outputs: same results
    d_lr = torch.from_numpy(lr).cuda()
    d_bits = torch.zeros((F, W), dtype=torch.int32, device="cuda")
    d_it = torch.zeros(F, dtype=torch.int32, device="cuda")
    d_ok = torch.zeros(F, dtype=torch.uint8, device="cuda")
    d_post = torch.zeros((F, N), dtype=torch.float64, device="cuda")
    dec.decode_device(ldpc.IN_LR_F64, d_lr.data_ptr(), F, 30, bits_ptr=d_bits.data_ptr(), iters_ptr=d_it.data_ptr(),
                      ok_ptr=d_ok.data_ptr(), post_ptr=d_post.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_it.cpu().numpy(), a["iters"]) and np.array_equal(d_ok.cpu().numpy(), a["ok"])
    assert np.array_equal(d_bits.cpu().numpy().view(np.uint32), a["bits_packed"])
    assert np.array_equal(d_post.cpu().numpy().view(np.uint64), a["post"].view(np.uint64))
    # pinned host buffers (asynchronous copies) and a strided input (every frame padded by 5 doubles)
    h_lr = torch.from_numpy(np.pad(lr, ((0, 0), (0, 5)))).pin_memory()
    h_bits = torch.zeros((F, W), dtype=torch.int32).pin_memory()
    h_it = torch.zeros(F, dtype=torch.int32).pin_memory()
    inp = ldpc.Input(kind=ldpc.IN_LR_F64, flags=0, data=h_lr.data_ptr(), frame_stride=(N + 5) * 8, param=0.0, table=None)
    out = ldpc.Output(bits=h_bits.data_ptr(), iters=h_it.data_ptr())
    assert ldpc.lib().dnaldpc_decode_batch(dec._h, C.byref(inp), F, 30, C.byref(out)) == 0, ldpc.lib().dnaldpc_last_error()
    assert np.array_equal(h_it.numpy(), a["iters"]) and np.array_equal(h_bits.numpy().view(np.uint32), a["bits_packed"])
    # LLR inputs exponentiated on the host (the CLI's path) are the same ratios
    llr = np.log(lr)
    b = dec.decode(ldpc.IN_LLR_F64, llr, 30, flags=ldpc.FLAG_HOST_EXP, want=("bits", "iters", "ok"))
    lr2 = np.exp(llr)
    for f in (0, 1500, F - 1):
        o = orc.decode(lr2[f], 30, want_post=False)
        assert b["iters"][f] == o["n"] and np.array_equal(b["bits"][f], o["dblk"]), f
    dec.close()


def test_decoders_for_two_codes_interleaved(code, orc, cws):
    """Two decoders for different codes on one GPU, used alternately: the shared-memory syndrome kernel's dynamic
    shared-memory limit is a per-function attribute that one decoder must not lower for the other."""
    row_ptr, col_idx = gen_regular_pchk.gen_regular(120, 60, 3, 1)
    small_code = ldpc.Code(csr=(60, 120, row_ptr, col_idx))
    small_orc = ol.Oracle(csr=(60, 120, row_ptr, col_idx))
    big = ldpc.Decoder(code, wave_frames=64)
    small = ldpc.Decoder(small_code, wave_frames=64)
    recv, lr, _ = _mixed_frames(cws, 40, 17)
    rs = np.random.RandomState(2)
    lr_s = np.exp(rs.randn(40, 120) * 2 + 1.5)
    for _ in range(2):
        a = big.decode(ldpc.IN_LR_F64, lr, 20)
        b = small.decode(ldpc.IN_LR_F64, lr_s, 20)
        for f in (0, 13, 39):
            o = orc.decode(lr[f], 20, want_post=False)
            assert a["iters"][f] == o["n"] and np.array_equal(a["bits"][f], o["dblk"])
            o = small_orc.decode(lr_s[f], 20, want_post=False)
            assert b["iters"][f] == o["n"] and np.array_equal(b["bits"][f], o["dblk"])
    big.close(); small.close()


def test_vote_table_on_the_device(code, cws):
    """A caller's VOTE_I8 table may live in host or device memory (dnaldpc_input.table)."""
    import torch
    rs = np.random.RandomState(8)
    F = 64
    reads = rs.poisson(3.9, (F, N)).astype(np.int16)
    k = reads - 2 * rs.binomial(reads, 0.02).astype(np.int16)
    k = np.where(cws[np.arange(F) % 272] == 0, k, -k).astype(np.int8)
    dec = ldpc.Decoder(code, wave_frames=64)
    a = dec.decode(ldpc.IN_VOTE_I8, k, 50, param=0.03)
    tab = ldpc.vote_table(0.03)
    b = dec.decode(ldpc.IN_VOTE_I8, k, 50, table=tab)
    d_k = torch.from_numpy(k).cuda()
    d_tab = torch.from_numpy(tab).cuda()
    d_bits = torch.zeros((F, W), dtype=torch.int32, device="cuda")
    d_it = torch.zeros(F, dtype=torch.int32, device="cuda")
    dec.decode_device(ldpc.IN_VOTE_I8, d_k.data_ptr(), F, 50, bits_ptr=d_bits.data_ptr(), iters_ptr=d_it.data_ptr(),
                      table_ptr=d_tab.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(a["iters"], b["iters"]) and np.array_equal(a["bits"], b["bits"])
    assert np.array_equal(d_it.cpu().numpy(), a["iters"]) and np.array_equal(d_bits.cpu().numpy().view(np.uint32), a["bits_packed"])
    dec.close()


def test_device_generators_match_host_twins(code, cws):
    """synth_vote_kernel is integer arithmetic: identical to oracle/bp_oracle.c:orc_synth_vote. synth_awgn_kernel does
    Box-Muller in fp64 with the device's log / sqrt / cos and rounds to float: equal to the libm twin except where the
    fp64 values straddle a float rounding boundary (a handful per million), and then by one ulp."""
    import torch
    F = 48
    cw_packed = np.packbits(cws.astype(np.uint8), axis=1, bitorder="little").view(np.uint32)
    d_cw = torch.from_numpy(cw_packed.astype(np.int32)).cuda()
    dec = ldpc.Decoder(code, wave_frames=64)
    st = torch.cuda.current_stream().cuda_stream
    d_k = torch.zeros((F, N), dtype=torch.int8, device="cuda")
    dec.synth_vote_device(d_cw.data_ptr(), 272, 5, 1000, F, 3.9, 0.01, d_k.data_ptr(), st)
    sigma = ldpc.std_dev(4.3, 1 - M / N)
    d_y = torch.zeros((F, N), dtype=torch.float32, device="cuda")
    dec.synth_awgn_device(d_cw.data_ptr(), 272, 5, 1000, F, sigma, d_y.data_ptr(), st)
    torch.cuda.synchronize()
    k, y = d_k.cpu().numpy(), d_y.cpu().numpy()
    for f in (0, 7, F - 1):
        assert np.array_equal(k[f], ol.synth_vote(cws[(1000 + f) % 272], 5, 1000 + f, N, 3.9, 0.01)), f
        yh = ol.synth_awgn(cws[(1000 + f) % 272], 5, 1000 + f, N, sigma)
        diff = y[f] != yh
        assert diff.mean() < 1e-3 and np.allclose(y[f], yh, rtol=3e-7, atol=0), (f, int(diff.sum()))
    # statistics: reads per bit ~ Poisson(3.9), noise ~ N(0, sigma^2)
    tx = np.stack([cws[(1000 + f) % 272] for f in range(F)])
    assert abs(np.abs(k).astype(np.float64).mean() - (3.9 * 0.98)) < 0.05
    noise = y - np.where(tx == 0, 1.0, -1.0)
    assert abs(noise.mean()) < 2e-3 and abs(noise.std() / sigma - 1) < 5e-3
    dec.close()


def test_redecode_sweep_resident_inputs(code, orc, cws):
    """dnaldpc_redecode_sweep_ex on vote counts (decoder.py:594-664): round r decodes the frames whose syndrome is still
    non-zero with eps_r; the int8 inputs stay in HBM. Checked frame by frame against the oracle run with the same
    per-round tables; a frame keeps the result of its last round."""
    rs = np.random.RandomState(12)
    F = 600
    reads = rs.poisson(3.9, (F, N)).astype(np.int16)
    wrong = rs.binomial(reads, 0.03).astype(np.int16)
    k = reads - 2 * wrong
    tx = cws[np.arange(F) % 272]
    k = np.where(tx == 0, k, -k).astype(np.int8)
    eps_rounds = [0.20, 0.10, 0.07, 0.05]  # oracle: no frame / about 60 % / most / all frames decode within 25 iterations
    dec = ldpc.Decoder(code, wave_frames=256)
    r = dec.redecode_sweep_ex(ldpc.IN_VOTE_I8, k, 25, eps_rounds, want=("bits", "iters", "ok", "post"))
    assert set(np.unique(r["rounds"])) - {0} and (r["rounds"] <= 3).all()  # several rounds really ran
    tabs = [ldpc.vote_table(e) for e in eps_rounds]
    # per-round single decodes give the expected round of every frame
    ok_prev = np.zeros(F, bool)
    exp_round = np.zeros(F, np.int32)
    for rd, e in enumerate(eps_rounds):
        todo = np.nonzero(~ok_prev)[0]
        if len(todo) == 0:
            break
        one = dec.decode(ldpc.IN_VOTE_I8, k[todo], 25, param=e)
        exp_round[todo] = rd
        ok_prev[todo] = one["ok"] == 1
    assert np.array_equal(r["rounds"], exp_round) and np.array_equal(r["ok"] == 1, ok_prev)
    for f in list(range(0, F, 97)) + list(np.nonzero(r["rounds"] == r["rounds"].max())[0][:3]):
        f = int(f)
        o = orc.decode(tabs[r["rounds"][f]][k[f].astype(np.int64) + 128], 25)
        assert r["iters"][f] == o["n"] and r["ok"][f] == o["ok"] and np.array_equal(r["bits"][f], o["dblk"]), f
        assert np.array_equal(r["post"][f].view(np.uint64), o["post"].view(np.uint64)), f
    # the LLR form of the same sweep (device exp) takes the same decisions
    L = np.array([np.log((1 - e) / e) for e in eps_rounds])
    llr = k.astype(np.float64) * L[0]
    s = dec.redecode_sweep_ex(ldpc.IN_LLR_F64, llr, 25, L / L[0])
    assert np.mean(s["rounds"] == r["rounds"]) > 0.97
    dec.close()


def _round_trip_1m(dec, code, cws, kind, gen, param, max_iter, pieces, piece_frames, check_frames, orc_lr):
    """1M frames in device-resident pieces: every frame flagged as a codeword must be the sent one; unflagged frames ran
    to max_iter; a sample against the oracle. Returns (frames, failures, frame-iterations)."""
    import torch
    st = torch.cuda.current_stream().cuda_stream
    cw_packed = np.packbits(cws.astype(np.uint8), axis=1, bitorder="little").view(np.uint32)
    d_cw = torch.from_numpy(cw_packed.astype(np.int32)).cuda()
    d_bits = torch.empty((piece_frames, W), dtype=torch.int32, device="cuda")
    d_it = torch.empty(piece_frames, dtype=torch.int32, device="cuda")
    d_ok = torch.empty(piece_frames, dtype=torch.uint8, device="cuda")
    fails = iters = 0
    for p in range(pieces):
        f0 = p * piece_frames
        d_in = gen(d_cw, f0, piece_frames, st)
        dec.decode_device(kind, d_in.data_ptr(), piece_frames, max_iter, param=param, bits_ptr=d_bits.data_ptr(),
                          iters_ptr=d_it.data_ptr(), ok_ptr=d_ok.data_ptr(), stream=st)
        torch.cuda.synchronize()
        okm = d_ok.bool()
        sent = d_cw[(torch.arange(piece_frames, device="cuda") + f0) % 272]
        assert torch.equal(d_bits[okm], sent[okm]), p
        assert bool((d_it[~okm] == max_iter).all()) and bool((d_it >= 0).all()) and bool((d_it <= max_iter).all())
        assert dec.stats()["frames"] == piece_frames and dec.stats()["frame_iters"] == int(d_it.sum())
        fails += int((~okm).sum())
        iters += int(d_it.sum())
        if p in (0, pieces - 1):
            it, bits, x = d_it.cpu().numpy(), d_bits.cpu().numpy().view(np.uint32), d_in[:check_frames].cpu().numpy()
            for f in range(check_frames):
                o = orc_lr(x[f])
                assert it[f] == o["n"], (p, f)
                assert np.array_equal(bits[f], np.packbits(o["dblk"].astype(np.uint8), bitorder="little").view(np.uint32)), (p, f)
        del d_in
    return pieces * piece_frames, fails, iters


def test_config3_vote_counts_one_million_frames(code, orc, cws):
    """BASELINE configs[2] at its stated size: 1 000 000 frames of per-bit vote counts (device generator: Poisson(3.9)
    reads per bit, 1 % read errors; eps = 0.02) through a decoder over every visible GPU."""
    import torch
    dec = ldpc.Decoder(code, devices=list(range(torch.cuda.device_count())), wave_frames=4096)
    vt = ldpc.vote_table(0.02)

    def gen(d_cw, f0, F, st):
        d = torch.empty((F, N), dtype=torch.int8, device="cuda")
        dec.synth_vote_device(d_cw.data_ptr(), 272, 21, f0, F, 3.9, 0.01, d.data_ptr(), st)
        return d
    frames, fails, iters = _round_trip_1m(dec, code, cws, ldpc.IN_VOTE_I8, gen, 0.02, 50, 4, 250000, 3,
                                          lambda x: orc.decode(vt[x.astype(np.int64) + 128], 50, want_post=False))
    assert frames == 1000000 and fails == 0 and 1.0 <= iters / frames <= 6.0, (fails, iters / frames)
    dec.close()


def test_config4_awgn_one_million_frames(code, orc, cws):
    """BASELINE configs[3] at its stated size: 1 000 000 AWGN frames (device generator, Eb/N0 = 4.6 dB, sigma =
    getStd_dev(4.6, 1 - M/N)) sharded over every visible GPU by the decoder itself (chunks pulled from one counter, the
    other GPUs read the resident floats through peer access). The device's exp() is outside the bit-exact boundary, so the
    sample is compared with the oracle at the decision level (same iteration count on nearly all frames, same bits)."""
    import torch
    dec = ldpc.Decoder(code, devices=list(range(torch.cuda.device_count())), wave_frames=4096)
    sigma = ldpc.std_dev(4.6, 1 - M / N)

    def gen(d_cw, f0, F, st):
        d = torch.empty((F, N), dtype=torch.float32, device="cuda")
        dec.synth_awgn_device(d_cw.data_ptr(), 272, 22, f0, F, sigma, d.data_ptr(), st)
        return d
    st = torch.cuda.current_stream().cuda_stream
    cw_packed = np.packbits(cws.astype(np.uint8), axis=1, bitorder="little").view(np.uint32)
    d_cw = torch.from_numpy(cw_packed.astype(np.int32)).cuda()
    PF, pieces = 62500, 16
    d_bits = torch.empty((PF, W), dtype=torch.int32, device="cuda")
    d_it = torch.empty(PF, dtype=torch.int32, device="cuda")
    d_ok = torch.empty(PF, dtype=torch.uint8, device="cuda")
    fails = iters = 0
    for p in range(pieces):
        d_in = gen(d_cw, p * PF, PF, st)
        dec.decode_device(ldpc.IN_AWGN_F32, d_in.data_ptr(), PF, 50, param=sigma, bits_ptr=d_bits.data_ptr(),
                          iters_ptr=d_it.data_ptr(), ok_ptr=d_ok.data_ptr(), stream=st)
        torch.cuda.synchronize()
        okm = d_ok.bool()
        sent = d_cw[(torch.arange(PF, device="cuda") + p * PF) % 272]
        assert torch.equal(d_bits[okm], sent[okm]), p
        assert bool((d_it[~okm] == 50).all())
        fails += int((~okm).sum()); iters += int(d_it.sum())
        if p == 0:
            it, y = d_it.cpu().numpy(), d_in[:6].cpu().numpy()
            same = 0
            for f in range(6):
                o = orc.decode(np.exp(2.0 * y[f].astype(np.float64) / (sigma * sigma)), 50, want_post=False)
                same += int(it[f] == o["n"])
                assert np.array_equal(d_bits[f].cpu().numpy().view(np.uint32),
                                      np.packbits(o["dblk"].astype(np.uint8), bitorder="little").view(np.uint32)), f
            assert same >= 5
        del d_in
    assert fails <= 5000 and 2.0 <= iters / (PF * pieces) <= 20.0, (fails, iters / (PF * pieces))
    dec.close()


def test_multi_device_dynamic_dispatch(code, cws):
    """The decoder's own multi-GPU dispatch (chunks pulled from one shared counter): host batches, device-resident
    batches (peer access) and the sweep give the results of a single GPU."""
    import torch
    nd = torch.cuda.device_count()
    if nd < 2:
        pytest.skip("needs at least 2 GPUs")
    F = 5000
    recv, lr, eps = _mixed_frames(cws, F, 77, (0.004, 0.006, 0.0075))
    packed = np.packbits(recv.astype(np.uint8), axis=1, bitorder="little").view(np.uint32)
    one = ldpc.Decoder(code, devices=[0], wave_frames=512)
    many = ldpc.Decoder(code, devices=list(range(nd)), wave_frames=512)
    a = one.decode(ldpc.IN_LR_F64, lr, 40, want=("bits", "iters", "ok", "post"))
    b = many.decode(ldpc.IN_LR_F64, lr, 40, want=("bits", "iters", "ok", "post"))
    for key in ("bits", "iters", "ok"):
        assert np.array_equal(a[key], b[key]), key
    assert np.array_equal(a["post"].view(np.uint64), b["post"].view(np.uint64))
    assert many.stats()["frames"] == F and many.stats()["frame_iters"] == int(a["iters"].sum())
    d_lr = torch.from_numpy(lr).cuda()
    d_bits = torch.zeros((F, W), dtype=torch.int32, device="cuda")
    d_it = torch.zeros(F, dtype=torch.int32, device="cuda")
    d_ok = torch.zeros(F, dtype=torch.uint8, device="cuda")
    many.decode_device(ldpc.IN_LR_F64, d_lr.data_ptr(), F, 40, bits_ptr=d_bits.data_ptr(), iters_ptr=d_it.data_ptr(),
                       ok_ptr=d_ok.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_it.cpu().numpy(), a["iters"]) and np.array_equal(d_ok.cpu().numpy(), a["ok"])
    assert np.array_equal(d_bits.cpu().numpy().view(np.uint32), a["bits_packed"])
    s1 = one.redecode_sweep_ex(ldpc.IN_BSC_BITS, packed, 10, [0.02, 0.01, 0.006])
    s2 = many.redecode_sweep_ex(ldpc.IN_BSC_BITS, packed, 10, [0.02, 0.01, 0.006])
    for key in ("bits", "iters", "ok", "rounds"):
        assert np.array_equal(s1[key], s2[key]), key
    one.close(); many.close()


def test_staging_rings_wrap_around(code, orc, cws):
    """Rings much shorter than the batch (test switches: 64-frame chunks, 2 chunks of input ring, 3 of output ring): every
    ring row is reused dozens of times while frames that run to max_iter hold the low-water mark back, so admission
    stalls on the output ring again and again. Results must not depend on any of it: compared with the default rings,
    with the device-resident path and with the oracle; also through two GPUs when there are two."""
    import torch
    F = 4000
    recv, lr, eps = _mixed_frames(cws, F, 123, (0.004, 0.006, 0.0075, 0.02))
    packed = np.packbits(recv.astype(np.uint8), axis=1, bitorder="little").view(np.uint32)
    t = ldpc.bsc_table(0.006)
    dec = ldpc.Decoder(code, wave_frames=256)
    ref = dec.decode(ldpc.IN_BSC_BITS, packed, 40, param=0.006, want=("bits", "dblk", "iters", "ok", "pchk"))
    os.environ["DNALDPC_CHUNK_FRAMES"] = "64"
    os.environ["DNALDPC_RING_CHUNKS"] = "2,3"
    try:
        a = dec.decode(ldpc.IN_BSC_BITS, packed, 40, param=0.006, want=("bits", "dblk", "iters", "ok", "pchk"))
        b = dec.decode(ldpc.IN_LR_F64, t[recv], 40, want=("bits", "iters", "ok", "post"))
        many = None
        if torch.cuda.device_count() > 1:
            d2 = ldpc.Decoder(code, devices=[0, 1], wave_frames=256)
            many = d2.decode(ldpc.IN_BSC_BITS, packed, 40, param=0.006, want=("bits", "iters", "ok"))
            d2.close()
    finally:
        del os.environ["DNALDPC_CHUNK_FRAMES"], os.environ["DNALDPC_RING_CHUNKS"]
    for key in ("bits", "dblk", "iters", "ok", "pchk"):
        assert np.array_equal(a[key], ref[key]), key
    for key in ("bits", "iters", "ok"):
        assert np.array_equal(b[key], ref[key]), key
        if many is not None:
            assert np.array_equal(many[key], ref[key]), key
    assert (ref["iters"] == 40).sum() > 100 and (ref["iters"] < 10).sum() > 1000  # stragglers and fast frames interleaved
    for f in list(range(0, F, 499)) + [int(np.argmax(ref["iters"]))]:
        o = orc.decode(t[recv[f]], 40)
        assert a["iters"][f] == o["n"] and np.array_equal(a["bits"][f], o["dblk"]), f
        assert np.array_equal(b["post"][f].view(np.uint64), o["post"].view(np.uint64)), f
    dec.close()
