"""GPU parity tests (run on the B200 box: pytest -m gpu). Every decode goes through the C ABI (libdnaldpc.so via
ctypes) and is compared with (a) the golden vectors produced by the unmodified reference and (b) the CPU oracle
(oracle/liboracle.so) on the same seeded inputs. fp64 bar: decoded bits, iteration counts, success flags, syndromes
AND posteriors bit-exact (tolerance 0; north_star allows 1e-9 relative on posteriors)."""
import hashlib
import os

import numpy as np
import pytest

import _pkg
import gen_regular_pchk
import oraclelib as ol

pytestmark = pytest.mark.gpu
ldpc = _pkg.load()


@pytest.fixture(scope="module")
def code18432():
    return ldpc.Code(ol.PCHK_18432)


@pytest.fixture(scope="module")
def dec18432(code18432):
    d = ldpc.Decoder(code18432, wave_frames=256)
    yield d
    d.close()


@pytest.fixture(scope="module")
def orc18432():
    return ol.Oracle(ol.PCHK_18432)


@pytest.fixture(scope="module")
def cws():
    return ol.load_codewords()


def _sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def _golden_check(dec, g, N):
    names = [str(n) for n in g["names"]]
    by_iter = {}
    for n in names:
        by_iter.setdefault(int(g[n + ".max_iter"]), []).append(n)
    for mi, group in by_iter.items():  # one batch per max_iter: frames of very different lengths share a warp group
        lr = np.stack([g[n + ".lratio"] for n in group])
        r = dec.decode(ldpc.IN_LR_F64, lr, mi, want=("bits", "dblk", "iters", "ok", "post", "pchk"))
        for k, n in enumerate(group):
            assert r["iters"][k] == int(g[n + ".n"]), n
            assert r["ok"][k] == int(g[n + ".ok"]), n
            assert np.array_equal(np.packbits(r["bits"][k].astype(np.uint8), bitorder="little"), g[n + ".dblk"]), n
            assert np.array_equal(r["dblk"][k], r["bits"][k]), n
            assert np.array_equal(np.packbits(r["pchk"][k], bitorder="little"), g[n + ".pchk"]), n
            assert np.array_equal(_sha(r["post"][k]), g[n + ".post_sha"]), n
            if n + ".post" in g:
                assert np.array_equal(r["post"][k].view(np.uint64), g[n + ".post"].view(np.uint64)), n


def test_math_sequences_match_ieee_division():
    """The inlined MUFU.RCP64H + Newton sequences equal nvcc's full-range IEEE division on 2^26 operands."""
    assert ldpc.selftest_math(1 << 26, seed=12345) == 0


def test_golden_n18432(dec18432):
    _golden_check(dec18432, np.load(os.path.join(ol.GOLDEN, "golden_n18432.npz")), 18432)


def test_golden_small():
    code = ldpc.Code(os.path.join(ol.GOLDEN, "small_n120_m60.pchk"))
    dec = ldpc.Decoder(code)
    _golden_check(dec, np.load(os.path.join(ol.GOLDEN, "golden_small.npz")), 120)
    dec.close()


def test_golden_n65536(tmp_path):
    g = np.load(os.path.join(ol.GOLDEN, "golden_n65536.npz"))
    row_ptr, col_idx = gen_regular_pchk.gen_regular(65536, 6554, 3, 5)
    path = str(tmp_path / "big.pchk")
    gen_regular_pchk.write_pchk(path, 6554, 65536, row_ptr, col_idx)
    assert hashlib.sha256(open(path, "rb").read()).hexdigest() == str(g["pchk_sha256"])
    dec = ldpc.Decoder(ldpc.Code(path))
    _golden_check(dec, g, 65536)
    dec.close()


def _bsc_lr(cw, flips, p):
    return np.where((cw ^ flips) == 0, (1 - p) / p, p / (1 - p)).astype(np.float64)


def test_batch_vs_oracle_mixed_convergence(code18432, orc18432, cws):
    """70 frames (ragged last group, 3 waves of 32) whose iteration counts differ inside one warp group."""
    N = 18432
    F = 70
    eps_list = [0.004, 0.006, 0.0075, 0.0085, 0.02]
    lr = np.zeros((F, N))
    for f in range(F):
        eps = eps_list[f % len(eps_list)]
        lr[f] = _bsc_lr(cws[f % 272], ol.bsc_flips(21, f, N, eps), eps)
    lr[5] = _bsc_lr(cws[5], np.zeros(N, np.int8), 0.02)  # noiseless: n == 0
    want = ("bits", "iters", "ok", "post", "pchk")
    for wave in (32, 64, 4096):
        dec = ldpc.Decoder(code18432, wave_frames=wave)
        r = dec.decode(ldpc.IN_LR_F64, lr, 30, want=want)
        dec.close()
        for f in range(F):
            o = orc18432.decode(lr[f], 30)
            assert r["iters"][f] == o["n"] and r["ok"][f] == o["ok"], (wave, f)
            assert np.array_equal(r["bits"][f], o["dblk"]), (wave, f)
            assert np.array_equal(r["pchk"][f].astype(np.int8), o["pchk"]), (wave, f)
            assert np.array_equal(r["post"][f].view(np.uint64), o["post"].view(np.uint64)), (wave, f)
        assert r["iters"][5] == 0 and r["ok"][5] == 1
    assert len(set(r["iters"][:32].tolist())) > 3  # the mask logic really was exercised


def test_input_kinds(dec18432, orc18432, cws):
    N = 18432
    rs = np.random.RandomState(77)
    F = 12
    # BSC packed bits == LR_F64 with the two-entry table
    p = 0.006
    recv = np.stack([cws[f] ^ ol.bsc_flips(5, f, N, p) for f in range(F)]).astype(np.uint8)
    packed = np.packbits(recv, axis=1, bitorder="little").view(np.uint32)
    a = dec18432.decode(ldpc.IN_BSC_BITS, packed, 50, param=p, want=("bits", "iters", "ok", "post"))
    t = ldpc.bsc_table(p)
    b = dec18432.decode(ldpc.IN_LR_F64, t[recv], 50, want=("bits", "iters", "ok", "post"))
    for k in ("bits", "iters", "ok"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["post"].view(np.uint64), b["post"].view(np.uint64))
    o = orc18432.decode(t[recv[3]], 50)
    assert a["iters"][3] == o["n"] and np.array_equal(a["bits"][3], o["dblk"])
    # vote counts (decoder.py:292-316): int8 k, LR = table[k+128]
    reads = rs.poisson(3.9, (F, N))
    wrong = rs.binomial(reads, 0.01)
    k = (reads - 2 * wrong)
    k = np.where(cws[:F] == 0, k, -k).astype(np.int8)
    vt = ldpc.vote_table(0.02)
    a = dec18432.decode(ldpc.IN_VOTE_I8, k, 100, param=0.02, want=("bits", "iters", "ok", "post"))
    for f in range(0, F, 5):
        o = orc18432.decode(vt[k[f].astype(np.int64) + 128], 100)
        assert a["iters"][f] == o["n"] and a["ok"][f] == o["ok"] and np.array_equal(a["bits"][f], o["dblk"])
        assert np.array_equal(a["post"][f].view(np.uint64), o["post"].view(np.uint64))
    assert a["ok"].all() and np.array_equal(a["bits"], cws[:F])
    # user-supplied table
    a2 = dec18432.decode(ldpc.IN_VOTE_I8, k, 100, table=vt, want=("bits", "iters"))
    assert np.array_equal(a2["bits"], a["bits"]) and np.array_equal(a2["iters"], a["iters"])
    # LLR with host exp == the reference's LDPC_Encode (exp by libm); device exp agrees to the last ulp or two
    llr = np.where(recv == 0, 1.0, -1.0) * np.log((1 - p) / p) * rs.uniform(0.6, 1.4, (F, N))
    lr = np.zeros_like(llr)
    ol.Oracle.lib().orc_lr_from_llr(llr.reshape(-1), llr.size, lr.reshape(-1))
    a = dec18432.decode(ldpc.IN_LLR_F64, llr, 50, flags=ldpc.FLAG_HOST_EXP, want=("bits", "iters", "ok", "post"))
    for f in (0, 7):
        o = orc18432.decode(lr[f], 50)
        assert a["iters"][f] == o["n"] and np.array_equal(a["bits"][f], o["dblk"])
        assert np.array_equal(a["post"][f].view(np.uint64), o["post"].view(np.uint64))
    d = dec18432.decode(ldpc.IN_LLR_F64, llr, 50, want=("bits", "iters", "ok", "post"))
    assert np.array_equal(d["ok"], a["ok"]) and np.array_equal(d["bits"][a["ok"] == 1], a["bits"][a["ok"] == 1])
    # AWGN: LLR = 2y/sigma^2 (channel.cpp:32); exp on the device, so compare at the decision level
    sigma = ldpc.std_dev(4.6, 1 - 2048 / 18432)
    y = np.where(cws[:F] == 0, 1.0, -1.0) + sigma * rs.randn(F, N)
    a64 = dec18432.decode(ldpc.IN_AWGN_F64, y, 100, param=sigma, want=("bits", "iters", "ok"))
    ref_lr = np.exp(2.0 * y / (sigma * sigma))
    n_ref = np.array([orc18432.decode(ref_lr[f], 100)["n"] for f in range(F)])
    assert a64["ok"].all() and np.array_equal(a64["bits"], cws[:F])
    assert np.mean(a64["iters"] == n_ref) >= 0.9
    a32 = dec18432.decode(ldpc.IN_AWGN_F32, y.astype(np.float32), 100, param=sigma, want=("bits", "ok"))
    assert a32["ok"].all() and np.array_equal(a32["bits"], cws[:F])


def test_run_bp_decoder_dropin(dec18432, orc18432, cws):
    """Same buffer contract as Run_Belief_Propagation_Decoder(H, lratio, dblk, pchk, &bIsCodeword) -> n."""
    N = 18432
    for eps, mi in [(0.006, 200), (0.02, 7)]:
        lr = _bsc_lr(cws[9], ol.bsc_flips(7, 9, N, eps), eps)
        r = dec18432.run_bp_decoder(lr, mi)
        o = orc18432.decode(lr, mi)
        assert r["n"] == o["n"] and r["ok"] == o["ok"]
        assert np.array_equal(r["dblk"], o["dblk"]) and np.array_equal(r["pchk"], o["pchk"])


def test_edge_cases(dec18432, cws):
    N = 18432
    # empty batch
    r = dec18432.decode(ldpc.IN_LR_F64, np.zeros((0, N)), 10)
    assert r["iters"].shape == (0,)
    # max_iter == 0: decision is the channel hard decision, success only for codewords
    lr = np.stack([_bsc_lr(cws[0], np.zeros(N, np.int8), 0.02), _bsc_lr(cws[1], ol.bsc_flips(1, 1, N, 0.01), 0.01)])
    r = dec18432.decode(ldpc.IN_LR_F64, lr, 0)
    assert list(r["iters"]) == [0, 0] and list(r["ok"]) == [1, 0]
    assert np.array_equal(r["bits"], (lr < 1).astype(np.int8))
    # all-erased frame (LR == 1 everywhere): ties -> 0 at init (lratio < 1), zero syndrome, n == 0
    r = dec18432.decode(ldpc.IN_LR_F64, np.ones((1, N)), 5)
    assert r["iters"][0] == 0 and r["ok"][0] == 1 and not r["bits"].any()
    # invalid likelihood ratios (negative / NaN) take the IEEE slow path instead of corrupting neighbours
    lr = np.stack([_bsc_lr(cws[2], ol.bsc_flips(2, 2, N, 0.006), 0.006)] * 2)
    lr[1, 100] = -3.0
    lr[1, 200] = np.nan
    r = dec18432.decode(ldpc.IN_LR_F64, lr, 20)
    assert r["ok"][0] == 1 and np.array_equal(r["bits"][0], cws[2])
    # argument errors
    with pytest.raises(ldpc.LdpcError):
        dec18432.decode(ldpc.IN_BSC_BITS, np.zeros((1, 576), np.uint32), 5, param=0.0)
    with pytest.raises(ldpc.LdpcError):
        dec18432.decode(ldpc.IN_LR_F64, np.ones((1, N)), -1)


def test_invalid_inputs_match_oracle(dec18432, orc18432, cws):
    """Negative ratios are outside the model, but the arithmetic is still IEEE: the slow path must agree."""
    N = 18432
    lr = _bsc_lr(cws[3], ol.bsc_flips(3, 3, N, 0.006), 0.006)
    lr[::501] = -lr[::501]
    r = dec18432.decode(ldpc.IN_LR_F64, lr[None], 6, want=("bits", "iters", "ok"))
    o = orc18432.decode(lr, 6)
    assert r["iters"][0] == o["n"] and r["ok"][0] == o["ok"] and np.array_equal(r["bits"][0], o["dblk"])


def test_synth_generator_and_roundtrip(code18432, cws):
    """Device generator == numpy twin of the counter RNG; encode -> BSC -> decode returns the codewords (size-
    independent property used at full batch sizes), and results do not depend on the wave size."""
    import torch
    N, W = 18432, 576
    dec = ldpc.Decoder(code18432, wave_frames=1024)
    F = 2048 + 17
    cw_packed = np.packbits(cws.astype(np.uint8), axis=1, bitorder="little").view(np.uint32)
    d_cw = torch.from_numpy(cw_packed.astype(np.int32)).cuda()
    d_in = torch.empty((F, W), dtype=torch.int32, device="cuda")
    eps = 0.004
    dec.synth_bsc_device(d_cw.data_ptr(), 272, 99, 1000, F, eps, d_in.data_ptr())
    torch.cuda.synchronize()
    got = d_in.cpu().numpy().view(np.uint32)
    for f in (0, 271, 272, F - 1):
        recv = cws[(1000 + f) % 272] ^ ol.bsc_flips(99, 1000 + f, N, eps)
        assert np.array_equal(np.packbits(recv.astype(np.uint8), bitorder="little").view(np.uint32), got[f])
    d_bits = torch.empty((F, W), dtype=torch.int32, device="cuda")
    d_it = torch.empty(F, dtype=torch.int32, device="cuda")
    d_ok = torch.empty(F, dtype=torch.uint8, device="cuda")
    dec.decode_device(ldpc.IN_BSC_BITS, d_in.data_ptr(), F, 100, param=eps, bits_ptr=d_bits.data_ptr(),
                      iters_ptr=d_it.data_ptr(), ok_ptr=d_ok.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert bool(d_ok.all())
    want = cw_packed[(1000 + np.arange(F)) % 272]
    assert np.array_equal(d_bits.cpu().numpy().view(np.uint32), want)
    it = d_it.cpu().numpy()
    assert it.min() >= 1 and it.max() < 30
    # a different wave size gives identical iteration counts
    dec2 = ldpc.Decoder(code18432, wave_frames=96)
    r = dec2.decode(ldpc.IN_BSC_BITS, got[:200], 100, param=eps)
    assert np.array_equal(r["iters"], it[:200])
    dec.close(); dec2.close()


def test_multi_device_sharding(code18432, cws):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    N = 18432
    F = 100
    lr = np.stack([_bsc_lr(cws[f], ol.bsc_flips(4, f, N, 0.006), 0.006) for f in range(F)])
    one = ldpc.Decoder(code18432, devices=[0])
    two = ldpc.Decoder(code18432, devices=[0, 1])
    a = one.decode(ldpc.IN_LR_F64, lr, 50, want=("bits", "iters", "ok", "post"))
    b = two.decode(ldpc.IN_LR_F64, lr, 50, want=("bits", "iters", "ok", "post"))
    for k in ("bits", "iters", "ok"):
        assert np.array_equal(a[k], b[k])
    assert np.array_equal(a["post"].view(np.uint64), b["post"].view(np.uint64))
    one.close(); two.close()


def test_fixed_iterations_and_redecode_sweep(code18432, orc18432, cws):
    """fixed-iteration BP (Run_Belief_Propagation_Decoder_SAVE, dec.cpp:192-223) and the pipeline's
    eps-sweep re-decoding (decoder.py:594-664) as one call."""
    N = 18432
    dec = ldpc.Decoder(code18432, wave_frames=64)
    F = 40
    lr = np.stack([_bsc_lr(cws[f], ol.bsc_flips(61, f, N, 0.006), 0.006) for f in range(F)])
    r = dec.decode(ldpc.IN_LR_F64, lr, 7, flags=ldpc.FLAG_FIXED_ITERS, want=("bits", "iters", "ok", "pchk"))
    assert (r["iters"] == 7).all()
    for f in (0, 13, F - 1):
        o = orc18432.decode_fixed(lr[f], 7)
        assert r["ok"][f] == o["ok"] and np.array_equal(r["bits"][f], o["dblk"]) and np.array_equal(r["pchk"][f].astype(np.int8), o["pchk"])
    # re-decode sweep: vote-count LLRs computed with a too pessimistic eps fail at first, then decode once rescaled
    rs = np.random.RandomState(8)
    F = 24
    eps = 0.02
    reads = rs.poisson(2.2, (F, N))
    k = reads - 2 * rs.binomial(reads, 0.04)
    tx = cws[np.arange(F) % 272]
    llr = np.where(tx == 0, k, -k) * np.log((1 - eps) / eps)
    e2 = eps - 0.0005 * np.arange(0, 12)
    scales = np.log((1 - e2) / e2) / np.log((1 - eps) / eps)       # decoder.py:611
    res = dec.redecode_sweep(llr, 60, scales, flags=ldpc.FLAG_HOST_EXP)
    # same thing by hand through the oracle: a frame moves to the next scale while its syndrome is non-zero
    for f in range(0, F, 5):
        for rnd, s in enumerate(scales):
            lr_f = np.zeros(N)
            ol.Oracle.lib().orc_lr_from_llr(np.ascontiguousarray(s * llr[f]), N, lr_f)
            o = orc18432.decode(lr_f, 60, want_post=False)
            if o["ok"] or rnd == len(scales) - 1:
                break
        assert res["rounds"][f] == rnd and res["ok"][f] == o["ok"] and res["iters"][f] == o["n"], f
        assert np.array_equal(res["bits"][f], o["dblk"]), f
    dec.close()


def test_minsum_decoder(code18432, orc18432, cws):
    """floating-point min-sum (Run_MSA_Decoder_INF, dec.cpp:1216-1250): decisions, n, syndromes and the
    posterior LLRs bit-exact against the oracle, for every input kind whose LLR is exact on the device."""
    N = 18432
    rs = np.random.RandomState(14)
    dec = ldpc.Decoder(code18432, wave_frames=64)
    F = 50
    eps_of = [0.0005, 0.0015, 0.0025, 0.003]   # min-sum decodes this (8,72) code only below eps ~ 0.003
    recv = np.stack([cws[f] ^ ol.bsc_flips(90, f, N, eps_of[f % 4]) for f in range(F)]).astype(np.uint8)
    llr = np.stack([np.where(recv[f] == 0, 1.0, -1.0) * np.log((1 - eps_of[f % 4]) / eps_of[f % 4]) for f in range(F)])
    llr[7] = np.where(cws[7] == 0, 1.0, -1.0) * rs.poisson(3.0, N) * np.log(49.0)   # exact zeros and many ties
    r = dec.decode(ldpc.IN_LLR_F64, llr, 40, flags=ldpc.FLAG_MINSUM, want=("bits", "iters", "ok", "post", "pchk"))
    for f in range(0, F, 3):
        o = orc18432.decode_minsum(llr[f], 40)
        assert r["iters"][f] == o["n"] and r["ok"][f] == o["ok"], f
        assert np.array_equal(r["bits"][f], o["dblk"]) and np.array_equal(r["pchk"][f].astype(np.int8), o["pchk"]), f
        assert np.array_equal(r["post"][f].view(np.uint64), o["post"].view(np.uint64)), f
    assert len(set(r["iters"].tolist())) >= 3
    # BSC bits: LLR table = log of the two ratios (channel.cpp:78,83)
    p = 0.002
    recv8 = np.stack([cws[f] ^ ol.bsc_flips(91, f, N, p) for f in range(8)]).astype(np.uint8)
    packed = np.packbits(recv8, axis=1, bitorder="little").view(np.uint32)
    a = dec.decode(ldpc.IN_BSC_BITS, packed, 40, param=p, flags=ldpc.FLAG_MINSUM)
    t = np.log(ldpc.bsc_table(p))
    for f in (0, 5):
        o = orc18432.decode_minsum(t[recv8[f]], 40)
        assert a["iters"][f] == o["n"] and np.array_equal(a["bits"][f], o["dblk"])
    # AWGN: LLR = 2y/sigma^2 is exact on the device (no exp in the LLR domain)
    sigma = ldpc.std_dev(5.0, 1 - 2048 / 18432)
    y = np.where(cws[:6] == 0, 1.0, -1.0) + sigma * rs.randn(6, N)
    a = dec.decode(ldpc.IN_AWGN_F64, y, 60, param=sigma, flags=ldpc.FLAG_MINSUM, want=("bits", "iters", "ok", "post"))
    for f in range(6):
        o = orc18432.decode_minsum(2.0 * y[f] / (sigma * sigma), 60)
        assert a["iters"][f] == o["n"] and a["ok"][f] == o["ok"] and np.array_equal(a["bits"][f], o["dblk"])
        assert np.array_equal(a["post"][f].view(np.uint64), o["post"].view(np.uint64))
    # vote counts: LLR = k * ln((1-eps)/eps)
    k = rs.poisson(3.5, (4, N)) - 2 * rs.binomial(4, 0.02, (4, N))
    k = np.where(cws[:4] == 0, k, -k).astype(np.int8)
    a = dec.decode(ldpc.IN_VOTE_I8, k, 60, param=0.02, flags=ldpc.FLAG_MINSUM)
    o = orc18432.decode_minsum(k[2].astype(np.float64) * np.log((1 - 0.02) / 0.02), 60)
    assert a["iters"][2] == o["n"] and np.array_equal(a["bits"][2], o["dblk"])
    dec.close()


def test_scheduler_fuzz_small_code():
    """Continuous batching under many shapes: every frame of every batch against the oracle (bits, n, flag, posterior),
    on the small (N=120) code where the oracle is cheap. Covers F < slots, F >> slots, ragged groups, frames that need
    no iteration, max_iter 0/1, and both algorithms."""
    path = os.path.join(ol.GOLDEN, "small_n120_m60.pchk")
    code = ldpc.Code(path)
    orc = ol.Oracle(path)
    rs = np.random.RandomState(2027)
    Nn = 120
    for wave, F, mi in [(32, 1, 20), (32, 33, 20), (64, 500, 30), (96, 97, 5), (32, 200, 0), (64, 130, 1), (128, 1000, 50)]:
        dec = ldpc.Decoder(code, wave_frames=wave)
        q = rs.choice([0.0, 0.02, 0.06, 0.10], size=F)              # per-frame channel quality; 0.0 = noiseless (n = 0)
        flips = rs.rand(F, Nn) < q[:, None]
        amp = rs.uniform(0.5, 6.0, (F, Nn))
        llr = np.where(flips, -1.0, 1.0) * amp                      # all-zero codeword
        llr[rs.rand(F, Nn) < 0.03] = 0.0                            # erasures: LR exactly 1, LLR exactly 0
        lr = np.exp(llr)
        r = dec.decode(ldpc.IN_LR_F64, lr, mi, want=("bits", "iters", "ok", "post", "pchk"))
        m = dec.decode(ldpc.IN_LLR_F64, llr, mi, flags=ldpc.FLAG_MINSUM, want=("bits", "iters", "ok", "post"))
        for f in range(F):
            o = orc.decode(lr[f], mi)
            assert r["iters"][f] == o["n"] and r["ok"][f] == o["ok"], (wave, F, mi, f)
            assert np.array_equal(r["bits"][f], o["dblk"]) and np.array_equal(r["pchk"][f].astype(np.int8), o["pchk"]), (wave, F, mi, f)
            assert np.array_equal(r["post"][f].view(np.uint64), o["post"].view(np.uint64)), (wave, F, mi, f)
            o = orc.decode_minsum(llr[f], mi)
            assert m["iters"][f] == o["n"] and m["ok"][f] == o["ok"] and np.array_equal(m["bits"][f], o["dblk"]), (wave, F, mi, f)
            assert np.array_equal(m["post"][f].view(np.uint64), o["post"].view(np.uint64)), (wave, F, mi, f)
        dec.close()


def test_drain_tail_compaction(code18432, orc18432, cws):
    """Drain-tail compaction (compact_plan_kernel / compact_move_kernel): once no frame is pending the stragglers are
    re-packed into the lowest slot groups. 700 frames through 512 slots, 1 frame in 8 far above the threshold so that it
    runs all 40 iterations while its neighbours finish in a handful: the batch must compact several times, results must
    be identical with compaction switched off, and the stragglers (the moved slots) must match the oracle bit for bit."""
    N, F, mi = 18432, 700, 40
    lr = np.zeros((F, N))
    # likelihood ratios per frame: most frames at eps 0.005, every 8th at 0.02 (never converges)
    for f in range(F):
        eps = 0.02 if f % 8 == 3 else 0.005
        lr[f] = _bsc_lr(cws[f % 272], ol.bsc_flips(77, f, N, eps), 0.01)
    want = ("bits", "iters", "ok", "post")
    dec = ldpc.Decoder(code18432, wave_frames=512)
    r = dec.decode(ldpc.IN_LR_F64, lr, mi, want=want)
    assert dec.stats()["compactions"] >= 2, dec.stats()
    os.environ["DNALDPC_NO_COMPACT"] = "1"
    try:
        r0 = dec.decode(ldpc.IN_LR_F64, lr, mi, want=want)
        assert dec.stats()["compactions"] == 0
    finally:
        del os.environ["DNALDPC_NO_COMPACT"]
    dec.close()
    for k in want:
        assert np.array_equal(r[k].view(np.uint64) if k == "post" else r[k], r0[k].view(np.uint64) if k == "post" else r0[k]), k
    assert (r["iters"] == mi).sum() >= F // 8 - 2 and (r["iters"] < 15).sum() > F // 2
    for f in list(range(3, F, 40)) + list(range(0, F, 97)):
        o = orc18432.decode(lr[f], mi)
        assert r["iters"][f] == o["n"] and r["ok"][f] == o["ok"] and np.array_equal(r["bits"][f], o["dblk"]), f
        assert np.array_equal(r["post"][f].view(np.uint64), o["post"].view(np.uint64)), f


@pytest.mark.parametrize("switches", [{}, {"DNALDPC_SW_ZERO": "1", "DNALDPC_SW_CHUNK": "64", "DNALDPC_SW_PIECE": "32"},
                                      {"DNALDPC_SW_THIN": "0", "DNALDPC_SW_PACKED": "0"}, {"DNALDPC_SW_LOCKSTEP": "1"}],
                         ids=["groups", "groups-zero-chunk64-piece32", "groups-no-straggler-mode", "lockstep"])
def test_sliding_window_decoder(switches, monkeypatch):
    """sliding-window BP for SC-LDPC codes (dnaldpc_decode_window = Run_SW_Decoder, dec.cpp:2092-2196) on the
    generated (3,6) SC code: (a) the golden vectors produced by the unmodified reference, (b) a ragged batch (frames that
    leave a window position after 0..max_iter extra updates share warp groups; 100 frames through 2 groups of slots)
    against the oracle, for several window sizes incl. win == L (one window over the whole code). Variants: continuous
    batching by group (the default), the same with every group's message arrays cleared per refill, 64-frame staging
    chunks and 32-frame copy pieces (groups wait for frames that are still on their way), the same without the straggler
    side arrays and lane packing, and the wave-lock-step schedule of the first version."""
    for k, v in switches.items():
        monkeypatch.setenv(k, v)
    g = np.load(os.path.join(ol.GOLDEN, "golden_sw.npz"))
    path = os.path.join(ol.GOLDEN, "sc_z32_l12.pchk")
    code = ldpc.Code(path)
    orc = ol.Oracle(path)
    Mv, Mc, L, w = g["Mv"], g["Mc"], int(g["L"]), int(g["w"])
    N = code.N
    dec = ldpc.Decoder(code, wave_frames=64)
    want = ("bits", "dblk", "iters", "ok", "pchk")
    for name in g["names"]:
        name = str(name)
        r = dec.decode_window(g[name + ".lratio"][None, :], int(g[name + ".max_iter"]), L, w, int(g[name + ".win"]), Mv, Mc, want=want)
        assert r["iters"][0] == int(g[name + ".n"]) and r["ok"][0] == int(g[name + ".ok"]), name
        assert np.array_equal(np.packbits(r["bits"][0].astype(np.uint8), bitorder="little"), g[name + ".dblk"]), name
        assert np.array_equal(r["dblk"][0], r["bits"][0]), name
        assert np.array_equal(np.packbits(r["pchk"][0], bitorder="little"), g[name + ".pchk"]), name
    rs = np.random.RandomState(4242)
    F = 100
    eps = rs.choice([0.0, 0.03, 0.05, 0.065, 0.08, 0.11], size=F)
    llr = np.where(rs.rand(F, N) < eps[:, None], -1.0, 1.0) * rs.uniform(1.0, 4.0, (F, N))
    llr[rs.rand(F, N) < 0.02] = 0.0
    lr = np.exp(llr)
    seen = set()
    for win, mi in [(3, 12), (4, 25), (12, 8), (5, 0)]:
        r = dec.decode_window(lr, mi, L, w, win, Mv, Mc, want=want)
        for f in range(F):
            o = orc.decode_sw(lr[f], mi, L, w, win, Mv, Mc)
            assert r["iters"][f] == o["n"] and r["ok"][f] == o["ok"], (win, mi, f)
            assert np.array_equal(r["bits"][f], o["dblk"]) and np.array_equal(r["pchk"][f].astype(np.int8), o["pchk"]), (win, mi, f)
            seen.add(o["n"])
    assert len(seen) > 5
    with pytest.raises(ldpc.LdpcError):  # node counts that do not match the matrix
        dec.decode_window(lr[:2], 5, L, w, 4, Mv * 2, Mc)
    dec.close()


def test_sliding_window_generic_degrees():
    """The window decoder on matrices that are not spatially coupled, with a window description laid over them: a
    column-weight-10 random code (row degree 20: the generic check / bit update loops instead of the register kernels)
    and the n=18432 code (row degree 72, column weight 8). Checks reach columns Init_SW_Decoder has not touched yet, so
    the zeros alloc_entry leaves in the messages are operands (the per-refill clearing path). Against the oracle."""
    import gen_regular_pchk
    rs = np.random.RandomState(808)
    row_ptr, col_idx = gen_regular_pchk.gen_regular(640, 320, 10, 3, no4cycle=False)
    cases = [((320, 640, row_ptr, col_idx), None, 8, 2, 3, [80] * 8 + [0], [36] * 8 + [32], 70, 6),
             (None, ol.PCHK_18432, 6, 2, 2, [3072] * 6 + [0], [340] * 6 + [8], 40, 4)]
    for csr, path, L, w, win, Mv, Mc, F, mi in cases:
        code = ldpc.Code(path) if path else ldpc.Code(csr=csr)
        orc = ol.Oracle(path) if path else ol.Oracle(csr=csr)
        Mv, Mc = np.array(Mv, np.int32), np.array(Mc, np.int32)
        eps = rs.choice([0.0, 0.004, 0.01, 0.03], size=F)
        lr = np.exp(np.where(rs.rand(F, code.N) < eps[:, None], -1.0, 1.0) * rs.uniform(2.0, 5.0, (F, code.N)))
        dec = ldpc.Decoder(code, wave_frames=32)
        r = dec.decode_window(lr, mi, L, w, win, Mv, Mc, want=("bits", "iters", "ok", "pchk"))
        for f in range(F):
            o = orc.decode_sw(lr[f], mi, L, w, win, Mv, Mc)
            assert r["iters"][f] == o["n"] and r["ok"][f] == o["ok"], (code.N, f)
            assert np.array_equal(r["bits"][f], o["dblk"]) and np.array_equal(r["pchk"][f].astype(np.int8), o["pchk"]), (code.N, f)
        dec.close()


def test_sliding_window_multi_device():
    """dnaldpc_decode_window with several devices: every device decodes a contiguous share of the frames; results equal
    the one-device decoder's (frames are independent, DNA_main.cpp:629-651)."""
    import torch
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs more than one GPU")
    g = np.load(os.path.join(ol.GOLDEN, "golden_sw.npz"))
    code = ldpc.Code(os.path.join(ol.GOLDEN, "sc_z32_l12.pchk"))
    Mv, Mc, L, w = g["Mv"], g["Mc"], int(g["L"]), int(g["w"])
    rs = np.random.RandomState(99)
    F = 64 * ndev + 37
    eps = rs.choice([0.0, 0.03, 0.05, 0.08], size=F)
    lr = np.exp(np.where(rs.rand(F, code.N) < eps[:, None], -1.0, 1.0) * rs.uniform(1.0, 4.0, (F, code.N)))
    one = ldpc.Decoder(code, devices=[0], wave_frames=64)
    many = ldpc.Decoder(code, devices=list(range(ndev)), wave_frames=64)
    want = ("bits", "dblk", "iters", "ok", "pchk")
    a = one.decode_window(lr, 10, L, w, 4, Mv, Mc, want=want)
    b = many.decode_window(lr, 10, L, w, 4, Mv, Mc, want=want)
    for k in want:
        assert np.array_equal(a[k], b[k]), k
    assert len(set(a["iters"].tolist())) > 2
    one.close(); many.close()


def test_compaction_stress_small_code():
    """Drain-tail compaction forced at almost every tick (threshold raised to 97 % of the packed region through the test
    switch): slots are moved again and again, next to slots that wait for their harvest, with posteriors and syndromes
    requested, for sum-product and min-sum. Every frame against the oracle on the small code."""
    path = os.path.join(ol.GOLDEN, "small_n120_m60.pchk")
    code = ldpc.Code(path)
    orc = ol.Oracle(path)
    rs = np.random.RandomState(77)
    Nn = 120
    os.environ["DNALDPC_COMPACT_PCT"] = "97"
    try:
        total = 0
        for wave, F, mi in [(256, 256, 40), (512, 700, 25), (128, 128, 60)]:
            dec = ldpc.Decoder(code, wave_frames=wave)
            q = rs.choice([0.02, 0.05, 0.08, 0.12], size=F)
            flips = rs.rand(F, Nn) < q[:, None]
            llr = np.where(flips, -1.0, 1.0) * rs.uniform(0.5, 5.0, (F, Nn))
            lr = np.exp(llr)
            r = dec.decode(ldpc.IN_LR_F64, lr, mi, want=("bits", "iters", "ok", "post", "pchk"))
            total += dec.stats()["compactions"]
            m = dec.decode(ldpc.IN_LLR_F64, llr, mi, flags=ldpc.FLAG_MINSUM, want=("bits", "iters", "ok", "post"))
            total += dec.stats()["compactions"]
            for f in range(F):
                o = orc.decode(lr[f], mi)
                assert r["iters"][f] == o["n"] and r["ok"][f] == o["ok"], (wave, F, f)
                assert np.array_equal(r["bits"][f], o["dblk"]) and np.array_equal(r["pchk"][f].astype(np.int8), o["pchk"]), (wave, F, f)
                assert np.array_equal(r["post"][f].view(np.uint64), o["post"].view(np.uint64)), (wave, F, f)
                o = orc.decode_minsum(llr[f], mi)
                assert m["iters"][f] == o["n"] and m["ok"][f] == o["ok"] and np.array_equal(m["bits"][f], o["dblk"]), (wave, F, f)
                assert np.array_equal(m["post"][f].view(np.uint64), o["post"].view(np.uint64)), (wave, F, f)
            dec.close()
        assert total >= 12, total
    finally:
        del os.environ["DNALDPC_COMPACT_PCT"]


def test_smem_check_kernel_mixed_groups_in_subprocess():
    """The shared-memory check kernel normally runs only in the steady state (full groups, nobody admitted). Its
    cp.async path for mixed groups (fresh / idle lanes) is forced here with the A/B switch, in a child process
    because the switch is read once per process, and compared frame by frame with the oracle."""
    import subprocess
    import sys
    code = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import _pkg, oraclelib as ol
ldpc = _pkg.load()
N = 18432
cws = ol.load_codewords(); orc = ol.Oracle(ol.PCHK_18432)
dec = ldpc.Decoder(ldpc.Code(ol.PCHK_18432), wave_frames=64)
F = 150
eps_of = [0.004, 0.006, 0.0075, 0.0085]
lr = np.stack([np.where((cws[f] ^ ol.bsc_flips(33, f, N, eps_of[f % 4])) == 0, (1 - eps_of[f % 4]) / eps_of[f % 4], eps_of[f % 4] / (1 - eps_of[f % 4])) for f in range(F)])
r = dec.decode(ldpc.IN_LR_F64, lr, 25, want=("bits", "iters", "ok", "post"))
for f in range(0, F, 4):
    o = orc.decode(lr[f], 25)
    assert r["iters"][f] == o["n"] and r["ok"][f] == o["ok"] and np.array_equal(r["bits"][f], o["dblk"]), f
    assert np.array_equal(r["post"][f].view(np.uint64), o["post"].view(np.uint64)), f
print("ok", sorted(set(r["iters"].tolist()))[:6])
'''
    env = dict(os.environ, DNALDPC_ROW_SMEM_ALWAYS="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True)
    assert res.returncode == 0 and res.stdout.startswith("ok"), res.stdout + res.stderr
