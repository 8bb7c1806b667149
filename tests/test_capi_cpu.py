"""CPU-side checks of the product library: it loads, exports every symbol include/dnaldpc.h declares, the host
logic (.pchk reader/writer, CSR/CSC tables, CheckRegular, channel tables) matches the oracle, and creating a decoder
without a GPU fails loudly (no CPU fallback). No compute calls here."""
import ctypes as C
import os
import re
import struct

import numpy as np
import pytest

import _pkg
import gen_regular_pchk
import oraclelib as ol

ldpc = _pkg.load()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_match_header():
    hdr = open(os.path.join(ROOT, "include", "dnaldpc.h")).read()
    declared = set(re.findall(r"\b(dnaldpc_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(ldpc.EXPORTS), declared ^ set(ldpc.EXPORTS)
    L = C.CDLL(ldpc.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in ldpc.lib().dnaldpc_version()


def test_pchk_reader_matches_oracle(tmp_path):
    code = ldpc.Code(ol.PCHK_18432)
    orc = ol.Oracle(ol.PCHK_18432)
    assert (code.M, code.N, code.E) == (2048, 18432, 147456)
    rp, ci, cp, ce = code.export()
    assert np.array_equal(rp, orc.row_ptr) and np.array_equal(ci, orc.col_idx)
    assert np.array_equal(cp, orc.col_ptr) and np.array_equal(ce, orc.col_edge)
    assert code.check_regular() == orc.check_regular() == (8, 1, 72, 1)
    out = str(tmp_path / "rt.pchk")
    code.write_pchk(out)
    assert open(out, "rb").read() == open(ol.PCHK_18432, "rb").read()


def test_pchk_irregular_and_errors(tmp_path):
    p = str(tmp_path / "x.pchk")
    # unsorted rows + duplicate entry: merged and sorted like mod2sparse_insert
    open(p, "wb").write(struct.pack("<12i", 0x5080, 2, 4, -2, 4, 1, 1, -1, 3, 2, 3, 0))
    c = ldpc.Code(p)
    o = ol.Oracle(p)
    rp, ci, cp, ce = c.export()
    assert c.E == 4 and list(rp) == [0, 2, 4] and list(ci) == [1, 2, 0, 3]
    assert np.array_equal(cp, o.col_ptr) and np.array_equal(ce, o.col_edge)
    assert c.check_regular() == o.check_regular()
    for words, rc in [((0x5081, 2, 4, 0), 3), ((0x5080, 2, 4, -1, 5, 0), 3), ((0x5080, 2, 4, 1, 0), 3),
                      ((0x5080, 2, 4, -1, 1), 3), ((0x5080, 0, 4, 0), 3), ((0x5080, 2, 4, -3, 1, 0), 3)]:
        open(p, "wb").write(struct.pack("<%di" % len(words), *words))
        with pytest.raises(ldpc.LdpcError) as ei:
            ldpc.Code(p)
        assert ei.value.rc == rc
    with pytest.raises(ldpc.LdpcError) as ei:
        ldpc.Code(str(tmp_path / "missing.pchk"))
    assert ei.value.rc == 2 and "Can't open parity check file" in str(ei.value)


def test_from_csr_and_generated_code():
    row_ptr, col_idx = gen_regular_pchk.gen_regular(120, 60, 3, 1)
    c = ldpc.Code(csr=(60, 120, row_ptr, col_idx))
    o = ol.Oracle(csr=(60, 120, row_ptr, col_idx))
    rp, ci, cp, ce = c.export()
    assert np.array_equal(rp, o.row_ptr) and np.array_equal(ci, o.col_idx) and np.array_equal(ce, o.col_edge)
    with pytest.raises(ldpc.LdpcError):
        ldpc.Code(csr=(60, 120, row_ptr, col_idx + 1000))


def test_channel_helpers_match_oracle():
    L = ol.Oracle.lib()
    assert ldpc.std_dev(4.3, 1 - 2048 / 18432) == L.orc_std_dev(4.3, 1 - 2048 / 18432)
    t = ldpc.bsc_table(0.02)
    assert t[0] == L.orc_bsc_lr(0, 0.02) and t[1] == L.orc_bsc_lr(1, 0.02)
    v = ldpc.vote_table(0.02)
    llr = np.array([L.orc_vote_llr(k, 0.02) for k in range(-128, 128)])
    lr = np.zeros(256)
    L.orc_lr_from_llr(llr, 256, lr)
    assert np.array_equal(v, lr)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    code = ldpc.Code(ol.PCHK_18432)
    with pytest.raises(ldpc.LdpcError) as ei:
        ldpc.Decoder(code)
    assert ei.value.rc == 4 and "no CPU fallback" in str(ei.value)


def test_cli_argument_errors_without_gpu(tmp_path):
    """The CLI's argument / input-file errors come before any CUDA call: the reference's `argc error!` (DNA_main.cpp:494-502)
    for a wrong argument count, incl. the sliding-window block of decoder type 60 (:424-429), unsupported decoder types,
    an unreadable .pchk (rcode.cpp:60-64) and a missing node-count file."""
    import shutil
    import subprocess
    exe = os.path.join(ROOT, "dna-ldpc-codes_b200", "ldpc")
    d = str(tmp_path)
    shutil.copyfile(os.path.join(ol.GOLDEN, "sc_z32_l12.pchk"), os.path.join(d, "sc.pchk"))

    def run(*a):
        return subprocess.run([exe] + list(a), cwd=d, capture_output=True, text=True)
    r = run("0", "0", "0", "7", "200", "1", "cw", "soft", "sc", "0", "0", "0")            # one argument short
    assert r.returncode == 1 and "argc error!" in r.stderr
    r = run("0", "60", "0", "7", "20", "1", "cw", "soft", "sc", "0", "0", "0", "0")       # type 60 without its block
    assert r.returncode == 1 and "argc error!" in r.stderr
    r = run("0", "60", "0", "7", "20", "1", "cw", "soft", "sc", "0", "0", "0", "0", "0", "3", "12", "4")
    assert r.returncode == 1 and "Can't open node-count file: sc.txt" in r.stderr
    r = run("0", "50", "0", "7", "20", "1", "cw", "soft", "sc", "0", "0", "0", "0")
    assert r.returncode == 1 and "not supported" in r.stderr
    r = run("0", "0", "0", "7", "20", "1", "cw", "soft", "nosuch", "0", "0", "0", "0")
    assert r.returncode == 1 and "nosuch.pchk" in r.stderr


def test_resident_worker_protocol_without_gpu(tmp_path):
    """`ldpc --serve` and the client mode of the same binary: the command line, the working directory, stdout / stderr and
    the exit code travel over the Unix socket (error paths only here: nothing reaches a GPU); without a worker the client
    runs locally; --shutdown stops the worker and removes the socket."""
    import subprocess
    import time
    LDPC = os.path.join(ROOT, "dna-ldpc-codes_b200", "ldpc")
    tmp = str(tmp_path)
    sock = os.path.join(tmp, "w.sock")
    env = dict(os.environ, DNALDPC_SOCKET=sock)
    bad_pchk = [LDPC, "0", "0", "0", "7", "200", "1", "a", "b", "nope", "0", "0", "0", "0"]
    bad_argc = [LDPC, "0", "0", "0", "7", "200", "1", "a", "b", "nope", "0", "0", "0"]
    local = [subprocess.run(c, cwd=tmp, capture_output=True) for c in (bad_pchk, bad_argc)]
    assert local[0].returncode == 1 and b"Can't open parity check file: nope.pchk" in local[0].stderr
    assert local[1].returncode == 1 and b"argc error!" in local[1].stderr
    # no worker yet: the client falls back to a local run, --shutdown reports that nobody answers
    r = subprocess.run(bad_pchk, cwd=tmp, capture_output=True, env=env)
    assert (r.returncode, r.stdout, r.stderr) == (local[0].returncode, local[0].stdout, local[0].stderr)
    r = subprocess.run([LDPC, "--shutdown"], capture_output=True, env=env)
    assert r.returncode == 1 and b"no worker answers" in r.stderr
    worker = subprocess.Popen([LDPC, "--serve"], cwd="/", env=env, stderr=subprocess.PIPE)
    try:
        for _ in range(200):
            if os.path.exists(sock):
                break
            time.sleep(0.02)
        assert os.path.exists(sock)
        for c, ref in zip((bad_pchk, bad_argc), local):
            r = subprocess.run(c, cwd=tmp, capture_output=True, env=env)
            assert (r.returncode, r.stdout, r.stderr) == (ref.returncode, ref.stdout, ref.stderr)
    finally:
        r = subprocess.run([LDPC, "--shutdown"], capture_output=True, env=env)
        worker.wait(timeout=20)
    assert r.returncode == 0 and worker.returncode == 0 and not os.path.exists(sock)


def test_generator_twins_are_deterministic_and_plausible():
    """Host twins of the device input generators (oracle/bp_oracle.c): keyed by (seed, frame, bit) only, Poisson reads and
    Gaussian noise with the requested parameters, the codeword's sign applied."""
    N = 18432
    cw = (np.arange(N) % 3 == 0).astype(np.int8)
    k1, k2 = ol.synth_vote(cw, 5, 77, N, 3.9, 0.01), ol.synth_vote(cw, 5, 77, N, 3.9, 0.01)
    assert np.array_equal(k1, k2) and not np.array_equal(k1, ol.synth_vote(cw, 5, 78, N, 3.9, 0.01))
    k0 = ol.synth_vote(None, 5, 77, N, 3.9, 0.01)
    assert np.array_equal(k1, np.where(cw == 0, k0, -k0))
    reads = np.abs(k0.astype(np.int32))                       # 1 % wrong reads: |k| is the read count almost always
    assert abs(reads.mean() - 3.9 * 0.98) < 0.08 and (k0 >= 0).mean() > 0.97
    y = ol.synth_awgn(cw, 9, 3, N, 0.5)
    assert np.array_equal(y, ol.synth_awgn(cw, 9, 3, N, 0.5))
    noise = y - np.where(cw == 0, 1.0, -1.0)
    assert abs(noise.mean()) < 0.02 and abs(noise.std() - 0.5) < 0.02
    # the noise is Box-Muller on rng streams 4 / 5 exactly as specified in include/dnaldpc.h
    j = np.arange(8)
    u1 = ((ol.rng_u64(9, 3, j, 4) >> np.uint64(11)).astype(np.float64) + 1.0) / 9007199254740992.0
    u2 = (ol.rng_u64(9, 3, j, 5) >> np.uint64(11)).astype(np.float64) / 9007199254740992.0
    want = (np.where(cw[:8] == 0, 1.0, -1.0) + 0.5 * np.sqrt(-2.0 * np.log(u1)) * np.cos(6.283185307179586476925286766559 * u2)).astype(np.float32)
    assert np.allclose(y[:8], want, rtol=1e-6, atol=0)
