"""BASELINE.json configs 2-5 as parity-test cases on the GPU (pytest -m gpu), through the C ABI.
Exact comparisons against the CPU oracle on subsets the oracle finishes in seconds; size-independent properties
(encode -> channel -> decode round trip, wave-size independence, fp32-vs-fp64 frame-error rate) on the larger batches."""
import numpy as np
import pytest

import _pkg
import gen_regular_pchk
import oraclelib as ol

pytestmark = pytest.mark.gpu
ldpc = _pkg.load()
N = 18432


@pytest.fixture(scope="module")
def code():
    return ldpc.Code(ol.PCHK_18432)


@pytest.fixture(scope="module")
def orc():
    return ol.Oracle(ol.PCHK_18432)


@pytest.fixture(scope="module")
def cws():
    return ol.load_codewords()


def _bsc_packed(cws, seed, f0, F, eps):
    recv = np.stack([cws[(f0 + f) % 272] ^ ol.bsc_flips(seed, f0 + f, N, eps) for f in range(F)]).astype(np.uint8)
    return np.packbits(recv, axis=1, bitorder="little").view(np.uint32), recv


def test_config2_eps_sweep_iteration_histogram(code, orc, cws):
    """configs[1]: eps sweep. 0.01-0.04 never decode (FER = 1, n = max_iter); the waterfall values exercise
    early exit. Success flags and iteration counts must equal the reference's on the same frames."""
    dec = ldpc.Decoder(code, wave_frames=512)
    for eps, F, n_orc in [(0.01, 40, 2), (0.02, 40, 2), (0.04, 40, 2), (0.006, 600, 24), (0.0075, 300, 8), (0.0085, 200, 4)]:
        packed, recv = _bsc_packed(cws, 31, 0, F, eps)
        r = dec.decode(ldpc.IN_BSC_BITS, packed, 100, param=eps)
        if eps >= 0.01:
            assert not r["ok"].any() and (r["iters"] == 100).all()
            if eps == 0.02:
                assert np.array_equal(r["bits"], recv.astype(np.int8))  # decisions never move at 0.02
        t = ldpc.bsc_table(eps)
        idx = np.linspace(0, F - 1, n_orc).astype(int)
        for f in idx:  # exact per-frame parity on a subset
            o = orc.decode(t[recv[f]], 100, want_post=False)
            assert r["iters"][f] == o["n"] and r["ok"][f] == o["ok"] and np.array_equal(r["bits"][f], o["dblk"]), (eps, f)
        good = r["ok"] == 1
        assert np.array_equal(r["bits"][good], cws[np.arange(F)[good] % 272])  # converged frames = the sent codewords
    dec.close()


def test_config3_vote_count_soft_input(code, orc, cws):
    """configs[2]: per-bit soft information from aligned reads (decoder.py:292-316): Poisson(3.9) reads per bit, 1 % read
    errors, eps = 0.02, no read -> LLR 0; plus the Phred-style variant."""
    rs = np.random.RandomState(3)
    F = 4096
    reads = rs.poisson(3.9, (F, N)).astype(np.int16)
    wrong = rs.binomial(reads, 0.01).astype(np.int16)
    k = reads - 2 * wrong
    tx = cws[np.arange(F) % 272]
    k = np.where(tx == 0, k, -k).astype(np.int8)
    dec = ldpc.Decoder(code, wave_frames=1024)
    r = dec.decode(ldpc.IN_VOTE_I8, k, 200, param=0.02)
    assert r["ok"].all() and np.array_equal(r["bits"], tx)
    assert 1 <= r["iters"].min() and r["iters"].max() <= 12
    vt = ldpc.vote_table(0.02)
    for f in (0, 1777, F - 1):
        o = orc.decode(vt[k[f].astype(np.int64) + 128], 200, want_post=False)
        assert r["iters"][f] == o["n"] and np.array_equal(r["bits"][f], o["dblk"])
    # Phred variant: per-read error probability p = 10^(-Q/10) from the empirical quality histogram of 72000_RS_Q_0.txt
    q = np.array([34, 32, 33, 23, 31, 12, 27, 21, 38, 24])
    w = np.array([88.1, 3.6, 2.4, 1.6, 1.4, 1.3, 0.6, 0.6, 0.2, 0.2])
    w = w / w.sum()
    F2 = 256
    llr = np.zeros((F2, N))
    tx2 = cws[np.arange(F2) % 272]
    for _ in range(4):  # up to 4 reads per bit
        present = rs.rand(F2, N) < 0.92
        Q = rs.choice(q, size=(F2, N), p=w)
        p = 10.0 ** (-Q / 10.0)
        flip = rs.rand(F2, N) < p
        obs = tx2 ^ flip
        llr += np.where(present, np.where(obs == 0, 1.0, -1.0) * np.log((1 - p) / p), 0.0)
    lr = np.exp(llr)
    r2 = dec.decode(ldpc.IN_LR_F64, lr, 200, want=("bits", "iters", "ok", "post"))
    assert r2["ok"].all() and np.array_equal(r2["bits"], tx2)
    o = orc.decode(lr[5], 200)
    assert r2["iters"][5] == o["n"] and np.array_equal(r2["post"][5].view(np.uint64), o["post"].view(np.uint64))
    dec.close()


def test_config4_awgn_sharded(code, orc, cws):
    """configs[3]: AWGN soft channel, frames sharded over the visible GPUs (1, 2, 4 or 8) with no collective.
    sigma = getStd_dev(EbNo, 1 - M/N) (channel.cpp:9-16, DNA_main.cpp:605); LLR = 2y/sigma^2."""
    import torch
    rs = np.random.RandomState(4)
    F = 2048
    tx = cws[np.arange(F) % 272]
    ndev = torch.cuda.device_count()
    one = ldpc.Decoder(code, devices=[0], wave_frames=512)
    many = ldpc.Decoder(code, devices=list(range(ndev)), wave_frames=512) if ndev > 1 else None
    for ebno, min_ok in [(4.6, 0.999), (4.3, 0.97)]:
        sigma = ldpc.std_dev(ebno, 1 - 2048 / 18432)
        y = (np.where(tx == 0, 1.0, -1.0) + sigma * rs.randn(F, N)).astype(np.float32)
        a = one.decode(ldpc.IN_AWGN_F32, y, 100, param=sigma)
        assert a["ok"].mean() >= min_ok
        good = a["ok"] == 1
        assert np.array_equal(a["bits"][good], tx[good])
        if many is not None:
            b = many.decode(ldpc.IN_AWGN_F32, y, 100, param=sigma)
            for key in ("bits", "iters", "ok"):
                assert np.array_equal(a[key], b[key]), key
        # the same frames through the exact path (LR computed on the host in fp64) agree at the decision level
        lr = np.exp(2.0 * y[:8].astype(np.float64) / (sigma * sigma))
        n_ref = np.array([orc.decode(lr[f], 100, want_post=False)["n"] for f in range(8)])
        assert np.mean(a["iters"][:8] == n_ref) >= 0.75
    one.close()
    if many is not None:
        many.close()


def test_config5_large_random_code():
    """configs[4]: Neal-style random regular code N=65536, column weight 3, rate 0.9 (H tables 786 KB, messages 1.5 MB
    per frame, row degrees 27-30 -> the generic irregular-row kernel)."""
    row_ptr, col_idx = gen_regular_pchk.gen_regular(65536, 6554, 3, 5)
    code = ldpc.Code(csr=(6554, 65536, row_ptr, col_idx))
    orc = ol.Oracle(csr=(6554, 65536, row_ptr, col_idx))
    assert code.E == 196608 and code.check_regular() == orc.check_regular()
    dec = ldpc.Decoder(code, wave_frames=256)
    F, Nb = 300, 65536

    def eps_of(f):
        return [0.003, 0.005, 0.006][f % 3]
    recv = np.stack([ol.bsc_flips(9, f, Nb, eps_of(f)) for f in range(F)]).astype(np.uint8)  # all-zero codeword
    lr = np.stack([ldpc.bsc_table(eps_of(f))[recv[f]] for f in range(F)])
    r = dec.decode(ldpc.IN_LR_F64, lr, 60, want=("bits", "iters", "ok", "post"))
    for f in (0, 1, 2, 150, F - 1):
        o = orc.decode(lr[f], 60)
        assert r["iters"][f] == o["n"] and r["ok"][f] == o["ok"] and np.array_equal(r["bits"][f], o["dblk"]), f
        assert np.array_equal(r["post"][f].view(np.uint64), o["post"].view(np.uint64)), f
    good = r["ok"] == 1
    assert good.mean() > 0.9 and not r["bits"][good].any()
    dec.close()


def test_fp32_mode_matches_fer(code, cws):
    """Optional fp32 mode: frame-error rate equal to the fp64 (= reference-exact) decoder's within a 95 % confidence
    interval, at two operating points in the waterfall and on high-confidence vote-count inputs."""
    d64 = ldpc.Decoder(code, wave_frames=1024)
    d32 = ldpc.Decoder(code, wave_frames=1024, precision=ldpc.PREC_F32)
    for eps, F in [(0.0072, 1500), (0.0080, 1500)]:
        packed, _ = _bsc_packed(cws, 55, 0, F, eps)
        a = d64.decode(ldpc.IN_BSC_BITS, packed, 100, param=eps)
        b = d32.decode(ldpc.IN_BSC_BITS, packed, 100, param=eps)
        fa, fb = 1 - a["ok"].mean(), 1 - b["ok"].mean()
        half = 1.96 * np.sqrt((fa * (1 - fa) + fb * (1 - fb)) / F) + 1e-9  # two-proportion 95 % interval
        assert abs(fa - fb) <= half, (eps, fa, fb, half)
        assert 0.02 < fa < 0.98  # the operating point really is in the waterfall
        both = (a["ok"] == 1) & (b["ok"] == 1)
        assert np.array_equal(a["bits"][both], b["bits"][both])
    rs = np.random.RandomState(6)
    F = 512
    reads = rs.poisson(3.9, (F, N)).astype(np.int16)
    k = reads - 2 * rs.binomial(reads, 0.01).astype(np.int16)
    tx = cws[np.arange(F) % 272]
    k = np.where(tx == 0, k, -k).astype(np.int8)
    b = d32.decode(ldpc.IN_VOTE_I8, k, 100, param=0.02)   # LR up to e^58: clamped, must still decode everything
    assert b["ok"].all() and np.array_equal(b["bits"], tx)
    d64.close()
    d32.close()


def test_config2_full_size_properties(code, orc, cws):
    """configs[1] at its full size: a 100 000-frame batch on one GPU, device-resident, through properties that do not
    need the oracle on every frame: (1) encode -> BSC(eps) -> decode round trip: every frame flagged as a codeword equals
    the sent codeword (eps = 0.005, well below the waterfall, and eps = 0.0075 where a few per cent fail); (2) a frame is
    flagged exactly when its iteration count stayed below max_iter or its last syndrome was zero; (3) the result does not
    depend on how the frames flow through the slots: a second run with a quarter of the slots (so different refill,
    drain-tail compaction and kernel-variant decisions) gives identical bits, counts and flags; (4) a sample of frames,
    stragglers included, against the oracle."""
    import torch
    N, W, F, mi = 18432, 576, 100000, 60
    cw_packed = np.packbits(cws.astype(np.uint8), axis=1, bitorder="little").view(np.uint32)
    d_cw = torch.from_numpy(cw_packed.astype(np.int32)).cuda()
    st = torch.cuda.current_stream().cuda_stream
    big = ldpc.Decoder(code, wave_frames=4096)
    small = ldpc.Decoder(code, wave_frames=1024)
    for eps, seed in ((0.005, 31), (0.0075, 32)):
        d_in = torch.empty((F, W), dtype=torch.int32, device="cuda")
        big.synth_bsc_device(d_cw.data_ptr(), 272, seed, 0, F, eps, d_in.data_ptr(), st)
        outs = []
        for dec in (big, small):
            d_bits = torch.empty((F, W), dtype=torch.int32, device="cuda")
            d_it = torch.empty(F, dtype=torch.int32, device="cuda")
            d_ok = torch.empty(F, dtype=torch.uint8, device="cuda")
            dec.decode_device(ldpc.IN_BSC_BITS, d_in.data_ptr(), F, mi, param=eps, bits_ptr=d_bits.data_ptr(),
                              iters_ptr=d_it.data_ptr(), ok_ptr=d_ok.data_ptr(), stream=st)
            torch.cuda.synchronize()
            outs.append((d_bits, d_it, d_ok))
        (b0, i0, k0), (b1, i1, k1) = outs
        assert torch.equal(b0, b1) and torch.equal(i0, i1) and torch.equal(k0, k1), eps          # (3)
        sent = d_cw[torch.arange(F, device="cuda") % 272]
        okm = k0.bool()
        assert torch.equal(b0[okm], sent[okm]), eps                                             # (1)
        assert bool((i0[~okm] == mi).all()) and bool((i0 >= 1).all()) and bool((i0 <= mi).all())  # (2)
        fer = 1.0 - float(okm.float().mean())
        assert (fer < 1e-3) if eps == 0.005 else (0.005 < fer < 0.5), (eps, fer)
        it = i0.cpu().numpy()
        bits = b0.cpu().numpy().view(np.uint32)
        sample = list(range(0, F, 9973)) + list(np.nonzero(it == mi)[0][:3]) + [int(np.argmax(np.where(it < mi, it, 0)))]
        for f in sample:                                                                         # (4)
            f = int(f)
            e = eps
            lr = np.where((cws[f % 272] ^ ol.bsc_flips(seed, f, N, eps)) == 0, (1 - e) / e, e / (1 - e))
            o = orc.decode(lr, mi, want_post=False)
            assert it[f] == o["n"] and int(k0[f]) == o["ok"], (eps, f)
            assert np.array_equal(bits[f], np.packbits(o["dblk"].astype(np.uint8), bitorder="little").view(np.uint32)), (eps, f)
    big.close(); small.close()
