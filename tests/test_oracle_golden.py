"""The CPU restatement (oracle/bp_oracle.c) against golden vectors produced by the UNMODIFIED reference
(tools/make_golden.py ran Run_Belief_Propagation_Decoder, dec.cpp:583-605, through oracle/_ref). CPU only."""
import hashlib
import os

import numpy as np
import pytest

import gen_regular_pchk
import oraclelib as ol


def _sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def _check_file(orc, g):
    for name in g["names"]:
        name = str(name)
        r = orc.decode(g[name + ".lratio"], int(g[name + ".max_iter"]), want_post=True, want_msgs=True)
        assert r["n"] == int(g[name + ".n"]), name
        assert r["ok"] == int(g[name + ".ok"]), name
        assert np.array_equal(np.packbits(r["dblk"].astype(np.uint8), bitorder="little"), g[name + ".dblk"]), name
        assert np.array_equal(np.packbits(r["pchk"].astype(np.uint8), bitorder="little"), g[name + ".pchk"]), name
        # bit-exact posteriors and final messages (hashes of the raw float64 bytes)
        assert np.array_equal(_sha(r["post"]), g[name + ".post_sha"]), name
        assert np.array_equal(_sha(r["pr"]), g[name + ".pr_sha"]), name
        assert np.array_equal(_sha(r["lr"]), g[name + ".lr_sha"]), name
        if name + ".post" in g:
            assert np.array_equal(r["post"].view(np.uint64), g[name + ".post"].view(np.uint64)), name


def test_golden_n18432():
    orc = ol.Oracle(ol.PCHK_18432)
    assert (orc.M, orc.N, orc.E) == (2048, 18432, 147456)
    assert orc.check_regular() == (8, 1, 72, 1)
    _check_file(orc, np.load(os.path.join(ol.GOLDEN, "golden_n18432.npz")))


def test_golden_small():
    orc = ol.Oracle(os.path.join(ol.GOLDEN, "small_n120_m60.pchk"))
    _check_file(orc, np.load(os.path.join(ol.GOLDEN, "golden_small.npz")))


def test_golden_n65536(tmp_path):
    g = np.load(os.path.join(ol.GOLDEN, "golden_n65536.npz"))
    row_ptr, col_idx = gen_regular_pchk.gen_regular(65536, 6554, 3, 5)
    path = str(tmp_path / "big.pchk")
    gen_regular_pchk.write_pchk(path, 6554, 65536, row_ptr, col_idx)
    assert hashlib.sha256(open(path, "rb").read()).hexdigest() == str(g["pchk_sha256"])
    orc = ol.Oracle(path)
    assert orc.E == 196608
    _check_file(orc, g)


def test_codewords_satisfy_checks():
    """All 272 shipped codewords have zero syndrome and decode with n == 0."""
    orc = ol.Oracle(ol.PCHK_18432)
    cws = ol.load_codewords()
    for i in range(0, 272, 17):
        w, _ = orc.check(cws[i])
        assert w == 0
    r = orc.decode(np.where(cws[5] == 0, 49.0, 1 / 49.0), 100)
    assert r["n"] == 0 and r["ok"] == 1 and np.array_equal(r["dblk"], cws[5])


def test_pchk_roundtrip_and_errors(tmp_path):
    orc = ol.Oracle(ol.PCHK_18432)
    out = str(tmp_path / "rt.pchk")
    assert orc.write_pchk(out)
    assert open(out, "rb").read() == open(ol.PCHK_18432, "rb").read()
    # unsorted rows / duplicate entries are merged like mod2sparse_insert (mod2sparse.cpp:521-524)
    import struct
    p = str(tmp_path / "dup.pchk")
    open(p, "wb").write(struct.pack("<12i", 0x5080, 2, 4, -2, 4, 1, 1, -1, 3, 2, 3, 0))
    o2 = ol.Oracle(p)
    assert o2.E == 4 and list(o2.row_ptr) == [0, 2, 4] and list(o2.col_idx) == [1, 2, 0, 3]
    # errors: bad magic, column out of range, entry before a row marker, truncated
    for words, err in [((0x5081, 2, 4, 0), 2), ((0x5080, 2, 4, -1, 5, 0), 3), ((0x5080, 2, 4, 1, 0), 3),
                       ((0x5080, 2, 4, -1, 1), 3), ((0x5080, 0, 4, 0), 3)]:
        open(p, "wb").write(struct.pack("<%di" % len(words), *words))
        with pytest.raises(IOError):
            ol.Oracle(p)


def test_rng_twin():
    """numpy twin of the C counter RNG (shared specification with the device generators)."""
    L = ol.Oracle.lib()
    bits = np.arange(0, 5000, 37)
    v = ol.rng_u64(7, 123, bits, 2)
    for b, x in zip(bits, v):
        assert L.orc_rng_u64(7, 123, int(b), 2) == int(x)
    assert abs(ol.bsc_flips(7, 3, 18432, 0.02).mean() - 0.02) < 0.004


def test_golden_sliding_window():
    """Sliding-window BP: orc_sw_decode against vectors produced by the unmodified reference's
    Run_SW_Decoder (dec.cpp:2092-2196) on the generated SC-LDPC code (tools/make_golden_sw.py)."""
    import gen_sc_pchk
    g = np.load(os.path.join(ol.GOLDEN, "golden_sw.npz"))
    path = os.path.join(ol.GOLDEN, "sc_z32_l12.pchk")
    M, N, row_ptr, col_idx, Mv, Mc = gen_sc_pchk.gen_sc(32, 12, 11)  # the generator still produces the committed file
    orc = ol.Oracle(path)
    assert (orc.M, orc.N) == (M, N) and np.array_equal(orc.row_ptr, row_ptr) and np.array_equal(orc.col_idx, col_idx)
    assert np.array_equal(Mv, g["Mv"]) and np.array_equal(Mc, g["Mc"])
    L, w = int(g["L"]), int(g["w"])
    for name in g["names"]:
        name = str(name)
        r = orc.decode_sw(g[name + ".lratio"], int(g[name + ".max_iter"]), L, w, int(g[name + ".win"]), Mv, Mc, want_msgs=True)
        assert r["n"] == int(g[name + ".n"]) and r["ok"] == int(g[name + ".ok"]), name
        assert np.array_equal(np.packbits(r["dblk"].astype(np.uint8), bitorder="little"), g[name + ".dblk"]), name
        assert np.array_equal(np.packbits(r["pchk"].astype(np.uint8), bitorder="little"), g[name + ".pchk"]), name
        assert np.array_equal(_sha(r["pr"]), g[name + ".pr_sha"]) and np.array_equal(_sha(r["lr"]), g[name + ".lr_sha"]), name
