"""ctypes bindings for the CHECKERS under oracle/ (test infrastructure only).

* ``Oracle``  - oracle/liboracle.so, our plain-C restatement (oracle/bp_oracle.c).
* ``RefLib``  - oracle/_ref/libldpc_ref.so, the unmodified reference objects + oracle/ref_harness.cpp.
                Present only where it was built (this container; it travels to the GPU box prebuilt).

Nothing in the product imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
GOLDEN = os.path.join(ROOT, "tests", "golden")
PCHK_18432 = os.path.join(GOLDEN, "decode_n18432_m2048_final.pchk")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_cp = np.ctypeslib.ndpointer(dtype=np.int8, flags="C_CONTIGUOUS")


def build_oracle():
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    src = os.path.join(ORACLE_DIR, "bp_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


class _Code(C.Structure):
    _fields_ = [("M", C.c_int), ("N", C.c_int), ("E", C.c_int),
                ("row_ptr", C.POINTER(C.c_int)), ("col_idx", C.POINTER(C.c_int)),
                ("col_ptr", C.POINTER(C.c_int)), ("col_edge", C.POINTER(C.c_int))]


class Oracle:
    """The restated decoder (bp_oracle.c) on one code."""
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(build_oracle())
            L.orc_read_pchk.restype = C.POINTER(_Code)
            L.orc_read_pchk.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
            L.orc_code_from_csr.restype = C.POINTER(_Code)
            L.orc_code_from_csr.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _ip]
            L.orc_code_free.argtypes = [C.POINTER(_Code)]
            L.orc_write_pchk.argtypes = [C.c_char_p, C.POINTER(_Code)]
            L.orc_check_regular.argtypes = [C.POINTER(_Code)] + [C.POINTER(C.c_int)] * 4
            L.orc_check.argtypes = [C.POINTER(_Code), _cp, C.c_void_p]
            L.orc_bp_decode.argtypes = [C.POINTER(_Code), _dp, C.c_int, _cp, C.c_void_p, C.POINTER(C.c_int),
                                        C.c_void_p, C.c_void_p, C.c_void_p]
            L.orc_bp_decode_fixed.argtypes = [C.POINTER(_Code), _dp, C.c_int, _cp, C.c_void_p, C.POINTER(C.c_int)]
            L.orc_minsum_decode.argtypes = [C.POINTER(_Code), _dp, C.c_int, _cp, C.c_void_p, C.POINTER(C.c_int), C.c_void_p]
            L.orc_bp_decode_f32.argtypes = [C.POINTER(_Code), _fp, C.c_int, _cp, C.POINTER(C.c_int)]
            L.orc_sw_schedule.argtypes = [C.c_int] * 6 + [_ip, _ip, _ip]
            L.orc_sw_decode.argtypes = [C.POINTER(_Code), _dp] + [C.c_int] * 5 + [_ip, _ip, _cp, _cp, C.POINTER(C.c_int),
                                        C.c_void_p, C.c_void_p, C.c_void_p]
            L.orc_bp_decode_many.restype = C.c_long
            L.orc_bp_decode_many.argtypes = [C.POINTER(_Code), _dp, C.c_int, C.c_int, _cp, _ip, _ip]
            L.orc_std_dev.restype = C.c_double
            L.orc_std_dev.argtypes = [C.c_double, C.c_double]
            L.orc_awgn_llr.restype = C.c_double
            L.orc_awgn_llr.argtypes = [C.c_double, C.c_double]
            L.orc_bsc_lr.restype = C.c_double
            L.orc_bsc_lr.argtypes = [C.c_int, C.c_double]
            L.orc_vote_llr.restype = C.c_double
            L.orc_vote_llr.argtypes = [C.c_int, C.c_double]
            L.orc_lr_from_llr.argtypes = [_dp, C.c_int, _dp]
            L.orc_rng_u64.restype = C.c_uint64
            L.orc_rng_u64.argtypes = [C.c_uint64] * 4
            L.orc_rng_uniform.restype = C.c_double
            L.orc_rng_uniform.argtypes = [C.c_uint64] * 4
            L.orc_rng_normal.restype = C.c_double
            L.orc_rng_normal.argtypes = [C.c_uint64] * 4
            L.orc_synth_awgn.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_double, _fp]
            L.orc_synth_vote.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_double, C.c_double, _cp]
            cls._lib = L
        return cls._lib

    def __init__(self, pchk_path=None, csr=None):
        L = self.lib()
        if pchk_path is not None:
            err = C.c_int(0)
            self._c = L.orc_read_pchk(os.fsencode(pchk_path), C.byref(err))
            if not self._c:
                raise IOError("orc_read_pchk(%s) failed, err=%d" % (pchk_path, err.value))
        else:
            M, N, row_ptr, col_idx = csr
            row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int32)
            col_idx = np.ascontiguousarray(col_idx, dtype=np.int32)
            self._c = L.orc_code_from_csr(M, N, len(col_idx), row_ptr, col_idx)
        c = self._c.contents
        self.M, self.N, self.E = c.M, c.N, c.E
        self.row_ptr = np.ctypeslib.as_array(c.row_ptr, shape=(self.M + 1,)).copy()
        self.col_idx = np.ctypeslib.as_array(c.col_idx, shape=(max(self.E, 1),)).copy()[:self.E]
        self.col_ptr = np.ctypeslib.as_array(c.col_ptr, shape=(self.N + 1,)).copy()
        self.col_edge = np.ctypeslib.as_array(c.col_edge, shape=(max(self.E, 1),)).copy()[:self.E]

    def __del__(self):
        try:
            self.lib().orc_code_free(self._c)
        except Exception:
            pass

    def write_pchk(self, path):
        return self.lib().orc_write_pchk(os.fsencode(path), self._c)

    def check_regular(self):
        v = [C.c_int() for _ in range(4)]
        self.lib().orc_check_regular(self._c, *[C.byref(x) for x in v])
        return tuple(x.value for x in v)

    def check(self, dblk):
        dblk = np.ascontiguousarray(dblk, dtype=np.int8)
        pchk = np.zeros(self.M, dtype=np.int8)
        w = self.lib().orc_check(self._c, dblk, pchk.ctypes.data)
        return w, pchk

    def decode(self, lratio, max_iter, want_post=True, want_msgs=False):
        """-> dict(n, ok, dblk[N] int8, pchk[M] int8, post[N] f64 | None, pr/lr[E] | None)"""
        lratio = np.ascontiguousarray(lratio, dtype=np.float64)
        assert lratio.shape == (self.N,)
        dblk = np.zeros(self.N, dtype=np.int8)
        pchk = np.zeros(self.M, dtype=np.int8)
        post = np.zeros(self.N, dtype=np.float64) if want_post else None
        pr = np.zeros(self.E, dtype=np.float64) if want_msgs else None
        lr = np.zeros(self.E, dtype=np.float64) if want_msgs else None
        ok = C.c_int(0)
        n = self.lib().orc_bp_decode(self._c, lratio, max_iter, dblk, pchk.ctypes.data, C.byref(ok),
                                     post.ctypes.data if want_post else None,
                                     pr.ctypes.data if want_msgs else None,
                                     lr.ctypes.data if want_msgs else None)
        return dict(n=n, ok=ok.value, dblk=dblk, pchk=pchk, post=post, pr=pr, lr=lr)

    def decode_fixed(self, lratio, max_iter):
        lratio = np.ascontiguousarray(lratio, dtype=np.float64)
        dblk = np.zeros(self.N, dtype=np.int8)
        pchk = np.zeros(self.M, dtype=np.int8)
        ok = C.c_int(0)
        n = self.lib().orc_bp_decode_fixed(self._c, lratio, max_iter, dblk, pchk.ctypes.data, C.byref(ok))
        return dict(n=n, ok=ok.value, dblk=dblk, pchk=pchk)

    def decode_minsum(self, llr, max_iter):
        llr = np.ascontiguousarray(llr, dtype=np.float64)
        dblk = np.zeros(self.N, dtype=np.int8)
        pchk = np.zeros(self.M, dtype=np.int8)
        L = np.array(llr, dtype=np.float64)  # n == 0 leaves it untouched; we define it as the channel LLR then
        ok = C.c_int(0)
        n = self.lib().orc_minsum_decode(self._c, llr, max_iter, dblk, pchk.ctypes.data, C.byref(ok), L.ctypes.data)
        return dict(n=n, ok=ok.value, dblk=dblk, pchk=pchk, post=L)

    def decode_sw(self, lratio, max_iter, L, w, win, Mv, Mc, code_type=0, want_msgs=False, dblk_init=2):
        """Run_SW_Decoder restated (sliding-window BP, SC-LDPC codes). dblk starts filled with `dblk_init` like the
        reference's g_bit_stream_trans (DNA_main.cpp:664-666). -> dict(n, ok, dblk, pchk, iters_pos[L], pr/lr)"""
        lratio = np.ascontiguousarray(lratio, dtype=np.float64)
        Mv = np.ascontiguousarray(Mv, dtype=np.int32); Mc = np.ascontiguousarray(Mc, dtype=np.int32)
        dblk = np.full(self.N, dblk_init, dtype=np.int8)
        pchk = np.zeros(self.M, dtype=np.int8)
        ip = np.zeros(L, dtype=np.int32)
        pr = np.zeros(self.E, dtype=np.float64) if want_msgs else None
        lr = np.zeros(self.E, dtype=np.float64) if want_msgs else None
        ok = C.c_int(0)
        n = self.lib().orc_sw_decode(self._c, lratio, max_iter, code_type, L, w, win, Mv, Mc, dblk, pchk, C.byref(ok),
                                     ip.ctypes.data, pr.ctypes.data if want_msgs else None, lr.ctypes.data if want_msgs else None)
        return dict(n=n, ok=ok.value, dblk=dblk, pchk=pchk, iters_pos=ip, pr=pr, lr=lr)

    def sw_schedule(self, L, w, win, Mv, Mc, code_type=0):
        Mv = np.ascontiguousarray(Mv, dtype=np.int32); Mc = np.ascontiguousarray(Mc, dtype=np.int32)
        sched = np.zeros((L, 8), dtype=np.int32)
        self.lib().orc_sw_schedule(self.M, self.N, code_type, L, w, win, Mv, Mc, sched.reshape(-1))
        return sched

    def decode_f32(self, lratio, max_iter):
        lratio = np.ascontiguousarray(lratio, dtype=np.float32)
        dblk = np.zeros(self.N, dtype=np.int8)
        ok = C.c_int(0)
        n = self.lib().orc_bp_decode_f32(self._c, lratio, max_iter, dblk, C.byref(ok))
        return dict(n=n, ok=ok.value, dblk=dblk)

    def decode_many(self, lratio, max_iter):
        lratio = np.ascontiguousarray(lratio, dtype=np.float64)
        F = lratio.shape[0]
        dblk = np.zeros((F, self.N), dtype=np.int8)
        iters = np.zeros(F, dtype=np.int32)
        ok = np.zeros(F, dtype=np.int32)
        self.lib().orc_bp_decode_many(self._c, lratio, F, max_iter, dblk, iters, ok)
        return dict(n=iters, ok=ok, dblk=dblk)


def rng_u64(seed, frame, bits, stream=0):
    """Vectorised numpy twin of orc_rng_u64 (bits = integer array)."""
    M64 = np.uint64(0xFFFFFFFFFFFFFFFF)

    def mix(z):
        z = z.astype(np.uint64)
        z ^= z >> np.uint64(30); z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27); z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
        return z
    with np.errstate(over="ignore"):
        frame = np.asarray(frame, dtype=np.uint64)
        bits = np.asarray(bits, dtype=np.uint64)
        x = mix(np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + frame + np.uint64(0x632BE59BD9B4E019))
        x = mix(x ^ (bits * np.uint64(0xD6E8FEB86659FD93) + np.uint64(stream) * np.uint64(0xA0761D6478BD642F)
                     + np.uint64(0x2545F4914F6CDD1D)))
    return x & M64


def bsc_threshold(eps):
    """flip iff (u64 >> 11) < thr ; the same integer compare is used on the device."""
    return np.uint64(int(eps * 9007199254740992.0))


def bsc_flips(seed, frame, n, eps):
    return ((rng_u64(seed, frame, np.arange(n)) >> np.uint64(11)) < bsc_threshold(eps)).astype(np.int8)


def synth_awgn(cw, seed, frame, n, sigma):
    """Host twin of synth_awgn_kernel (oracle/bp_oracle.c:orc_synth_awgn): float32 [n] received values of one frame."""
    y = np.zeros(n, dtype=np.float32)
    cwp = None if cw is None else np.ascontiguousarray(cw, dtype=np.int8)
    Oracle.lib().orc_synth_awgn(None if cwp is None else cwp.ctypes.data, seed, frame, n, sigma, y)
    return y


def synth_vote(cw, seed, frame, n, mean_reads, read_err):
    """Host twin of synth_vote_kernel (oracle/bp_oracle.c:orc_synth_vote): int8 [n] vote counts of one frame."""
    k = np.zeros(n, dtype=np.int8)
    cwp = None if cw is None else np.ascontiguousarray(cw, dtype=np.int8)
    Oracle.lib().orc_synth_vote(None if cwp is None else cwp.ctypes.data, seed, frame, n, mean_reads, read_err, k)
    return k


class RefLib:
    """The unmodified reference decoder (oracle/_ref/libldpc_ref.so). One code per process (global state)."""
    path = os.path.join(ORACLE_DIR, "_ref", "libldpc_ref.so")
    _lib = None
    _loaded_pchk = None

    @classmethod
    def available(cls):
        return os.path.exists(cls.path)

    def __init__(self, pchk_path):
        cls = RefLib
        if cls._lib is None:
            L = C.CDLL(cls.path)
            L.ref_load_pchk.argtypes = [C.c_char_p]
            L.ref_export_csr.argtypes = [_ip, _ip]
            L.ref_check_regular.argtypes = [C.POINTER(C.c_int)] * 4
            L.ref_decode.argtypes = [_dp, C.c_int, _cp, _cp, C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_void_p]
            L.ref_check.argtypes = [_cp, _cp]
            L.ref_decode_fixed.argtypes = [_dp, C.c_int, _cp, _cp, C.POINTER(C.c_int)]
            L.ref_decode_minsum.argtypes = [_dp, C.c_int, _cp, _cp, C.POINTER(C.c_int), _dp]
            L.ref_decode_sw.argtypes = [_dp] + [C.c_int] * 5 + [_ip, _ip, _cp, _cp, C.POINTER(C.c_int), C.c_void_p, C.c_void_p]
            L.ref_decode_many.restype = C.c_long
            L.ref_decode_many.argtypes = [_dp, C.c_int, C.c_int, _cp, _ip, _ip]
            cls._lib = L
        if cls._loaded_pchk is not None and cls._loaded_pchk != pchk_path:
            raise RuntimeError("RefLib holds global state: one .pchk per process (use a subprocess for another)")
        if cls._loaded_pchk is None:
            if not os.path.exists(pchk_path):
                raise IOError(pchk_path)  # read_pchk would exit(1)
            cls._lib.ref_load_pchk(os.fsencode(pchk_path))
            cls._loaded_pchk = pchk_path
        L = cls._lib
        self.M, self.N, self.E = L.ref_M(), L.ref_N(), L.ref_E()

    def export_csr(self):
        row_ptr = np.zeros(self.M + 1, dtype=np.int32)
        col_idx = np.zeros(self.E, dtype=np.int32)
        self._lib.ref_export_csr(row_ptr, col_idx)
        return row_ptr, col_idx

    def check_regular(self):
        v = [C.c_int() for _ in range(4)]
        self._lib.ref_check_regular(*[C.byref(x) for x in v])
        return tuple(x.value for x in v)

    def check(self, dblk):
        dblk = np.ascontiguousarray(dblk, dtype=np.int8)
        pchk = np.zeros(self.M, dtype=np.int8)
        return self._lib.ref_check(dblk, pchk), pchk

    def decode(self, lratio, max_iter, want_post=True, want_msgs=False):
        lratio = np.ascontiguousarray(lratio, dtype=np.float64)
        dblk = np.zeros(self.N, dtype=np.int8)
        pchk = np.zeros(self.M, dtype=np.int8)
        post = np.zeros(self.N, dtype=np.float64) if want_post else None
        pr = np.zeros(self.E, dtype=np.float64) if want_msgs else None
        lr = np.zeros(self.E, dtype=np.float64) if want_msgs else None
        ok = C.c_int(0)
        n = self._lib.ref_decode(lratio, max_iter, dblk, pchk, C.byref(ok),
                                 post.ctypes.data if want_post else None,
                                 pr.ctypes.data if want_msgs else None,
                                 lr.ctypes.data if want_msgs else None)
        return dict(n=n, ok=ok.value, dblk=dblk, pchk=pchk, post=post, pr=pr, lr=lr)

    def decode_fixed(self, lratio, max_iter):
        lratio = np.ascontiguousarray(lratio, dtype=np.float64)
        dblk = np.zeros(self.N, dtype=np.int8)
        pchk = np.zeros(self.M, dtype=np.int8)
        ok = C.c_int(0)
        n = self._lib.ref_decode_fixed(lratio, max_iter, dblk, pchk, C.byref(ok))
        return dict(n=n, ok=ok.value, dblk=dblk, pchk=pchk)

    def decode_minsum(self, llr, max_iter):
        llr = np.ascontiguousarray(llr, dtype=np.float64)
        dblk = np.zeros(self.N, dtype=np.int8)
        pchk = np.zeros(self.M, dtype=np.int8)
        L = np.array(llr, dtype=np.float64)
        ok = C.c_int(0)
        n = self._lib.ref_decode_minsum(llr, max_iter, dblk, pchk, C.byref(ok), L)
        return dict(n=n, ok=ok.value, dblk=dblk, pchk=pchk, post=L)

    def decode_sw(self, lratio, max_iter, L, w, win, Mv, Mc, code_type=0, want_msgs=False, dblk_init=2):
        """The reference's Run_SW_Decoder (dec.cpp:2092-2196); dblk pre-filled with 2 like Alloc_Mem does."""
        lratio = np.ascontiguousarray(lratio, dtype=np.float64)
        Mv = np.ascontiguousarray(Mv, dtype=np.int32); Mc = np.ascontiguousarray(Mc, dtype=np.int32)
        dblk = np.full(self.N, dblk_init, dtype=np.int8)
        pchk = np.zeros(self.M, dtype=np.int8)
        pr = np.zeros(self.E, dtype=np.float64) if want_msgs else None
        lr = np.zeros(self.E, dtype=np.float64) if want_msgs else None
        ok = C.c_int(0)
        n = self._lib.ref_decode_sw(lratio, max_iter, code_type, L, w, win, Mv, Mc, dblk, pchk, C.byref(ok),
                                    pr.ctypes.data if want_msgs else None, lr.ctypes.data if want_msgs else None)
        return dict(n=n, ok=ok.value, dblk=dblk, pchk=pchk, pr=pr, lr=lr)

    def decode_many(self, lratio, max_iter):
        lratio = np.ascontiguousarray(lratio, dtype=np.float64)
        F = lratio.shape[0]
        dblk = np.zeros((F, self.N), dtype=np.int8)
        iters = np.zeros(F, dtype=np.int32)
        ok = np.zeros(F, dtype=np.int32)
        self._lib.ref_decode_many(lratio, F, max_iter, dblk, iters, ok)
        return dict(n=iters, ok=ok, dblk=dblk)


def load_codewords():
    """272 true codewords of the n=18432 code as int8 [272][18432] (fixture packed by tools/make_golden.py)."""
    packed = np.fromfile(os.path.join(GOLDEN, "codewords_n18432_272.bits"), dtype=np.uint8)
    return np.unpackbits(packed.reshape(272, -1), axis=1, bitorder="little")[:, :18432].astype(np.int8)
