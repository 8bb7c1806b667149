"""Pins the CPU restatement against the UNMODIFIED reference objects (oracle/_ref/libldpc_ref.so) on fresh
seeded inputs - beyond the committed golden vectors. Skipped where oracle/_ref was never built. CPU only."""
import numpy as np
import pytest

import oraclelib as ol

pytestmark = pytest.mark.skipif(not ol.RefLib.available(), reason="oracle/_ref not built (needs /root/reference)")


@pytest.fixture(scope="module")
def pair():
    return ol.RefLib(ol.PCHK_18432), ol.Oracle(ol.PCHK_18432)


def test_structure(pair):
    ref, orc = pair
    rp, ci = ref.export_csr()
    assert np.array_equal(rp, orc.row_ptr) and np.array_equal(ci, orc.col_idx)
    assert ref.check_regular() == orc.check_regular()


def test_bitexact_random_frames(pair):
    ref, orc = pair
    cws = ol.load_codewords()
    rs = np.random.RandomState(2024)
    N = ref.N
    for t in range(8):
        f = int(rs.randint(272))
        kind = t % 4
        if kind == 0:
            eps = [0.005, 0.0075][t // 4 % 2]
            lr = np.where((cws[f] ^ ol.bsc_flips(100 + t, f, N, eps)) == 0, (1 - eps) / eps, eps / (1 - eps))
        elif kind == 1:
            sigma = 1 / np.sqrt(2 * (1 - 2048 / 18432) * 10 ** 0.42)
            lr = np.exp(2 * (np.where(cws[f] == 0, 1.0, -1.0) + sigma * rs.randn(N)) / sigma ** 2)
        elif kind == 2:
            k = rs.poisson(3.7, N) - 2 * rs.binomial(4, 0.02, N)
            lr = np.exp(np.where(cws[f] == 0, k, -k) * np.log(49.0))
        else:
            lr = np.exp(rs.uniform(-60, 60, N))  # garbage: exercises inf / NaN guards
        mi = [100, 30, 100, 12][kind]
        a = ref.decode(lr, mi, want_msgs=True)
        b = orc.decode(lr, mi, want_msgs=True)
        assert a["n"] == b["n"] and a["ok"] == b["ok"]
        assert np.array_equal(a["dblk"], b["dblk"]) and np.array_equal(a["pchk"], b["pchk"])
        for key in ("post", "pr", "lr"):
            assert np.array_equal(a[key].view(np.uint64), b[key].view(np.uint64)), key


def test_check_matches(pair):
    ref, orc = pair
    rs = np.random.RandomState(5)
    d = (rs.rand(ref.N) < 0.3).astype(np.int8)
    wa, pa = ref.check(d)
    wb, pb = orc.check(d)
    assert wa == wb and np.array_equal(pa, pb)


def test_fixed_iteration_variant(pair):
    """Run_Belief_Propagation_Decoder_SAVE (dec.cpp:192-223): no early exit."""
    ref, orc = pair
    cws = ol.load_codewords()
    N = ref.N
    for f, eps, mi in [(3, 0.006, 9), (4, 0.0085, 5), (5, 0.004, 0)]:
        lr = np.where((cws[f] ^ ol.bsc_flips(77, f, N, eps)) == 0, (1 - eps) / eps, eps / (1 - eps))
        a, b = ref.decode_fixed(lr, mi), orc.decode_fixed(lr, mi)
        assert a["n"] == b["n"] == mi and a["ok"] == b["ok"]
        assert np.array_equal(a["dblk"], b["dblk"]) and np.array_equal(a["pchk"], b["pchk"])


def test_minsum_variant(pair):
    """Run_MSA_Decoder_INF (dec.cpp:1216-1250): floating-point min-sum in the LLR domain."""
    ref, orc = pair
    cws = ol.load_codewords()
    N = ref.N
    rs = np.random.RandomState(12)
    sigma = 1 / np.sqrt(2 * (1 - 2048 / 18432) * 10 ** 0.5)
    cases = []
    for f, eps, mi in [(3, 0.004, 50), (4, 0.006, 30), (5, 0.004, 0), (6, 0.02, 3)]:
        cases.append((np.where((cws[f] ^ ol.bsc_flips(78, f, N, eps)) == 0, 1.0, -1.0) * np.log((1 - eps) / eps), mi))
    cases.append((2 * (np.where(cws[7] == 0, 1.0, -1.0) + sigma * rs.randn(N)) / sigma ** 2, 40))
    llr = np.where(cws[8] == 0, 1.0, -1.0) * rs.poisson(3.0, N) * np.log(49.0)   # many exact zeros and ties
    cases.append((llr, 25))
    for llr, mi in cases:
        a, b = ref.decode_minsum(llr, mi), orc.decode_minsum(llr, mi)
        assert a["n"] == b["n"] and a["ok"] == b["ok"]
        assert np.array_equal(a["dblk"], b["dblk"]) and np.array_equal(a["pchk"], b["pchk"])
        assert np.array_equal(a["post"].view(np.uint64), b["post"].view(np.uint64))


def test_sliding_window_variant_in_subprocess():
    """orc_sw_decode vs the reference's Run_SW_Decoder (dec.cpp:2092-2196) on fresh seeded inputs, every final message
    included. Child process: the reference object holds one .pchk per process."""
    import os
    import subprocess
    import sys
    code = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.getcwd(), "tests")); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import oraclelib as ol, gen_sc_pchk
path = os.path.join(ol.GOLDEN, "sc_z32_l12.pchk")
M, N, rp, ci, Mv, Mc = gen_sc_pchk.gen_sc(32, 12, 11)
ref, orc = ol.RefLib(path), ol.Oracle(path)
rs = np.random.RandomState(99)
iters = set()
for t in range(24):
    eps = [0.03, 0.05, 0.065, 0.08][t % 4]
    win = [3, 4, 6, 12][(t // 4) % 4]
    mi = [1, 8, 25][t % 3]
    llr = np.where(rs.rand(N) < eps, -1.0, 1.0) * rs.uniform(1.0, 4.0, N)
    llr[rs.rand(N) < 0.02] = 0.0
    lr = np.exp(llr)
    a = ref.decode_sw(lr, mi, 12, 3, win, Mv, Mc, want_msgs=True)
    b = orc.decode_sw(lr, mi, 12, 3, win, Mv, Mc, want_msgs=True)
    assert a["n"] == b["n"] and a["ok"] == b["ok"], t
    assert np.array_equal(a["dblk"], b["dblk"]) and np.array_equal(a["pchk"], b["pchk"]), t
    assert np.array_equal(a["pr"].view(np.uint64), b["pr"].view(np.uint64)) and np.array_equal(a["lr"].view(np.uint64), b["lr"].view(np.uint64)), t
    iters.add(a["n"])
print("ok", sorted(iters))
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True)
    assert res.returncode == 0 and res.stdout.startswith("ok"), res.stdout + res.stderr
