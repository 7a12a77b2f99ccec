"""oracle/oracle.py -- TEST INFRASTRUCTURE.  ctypes binding of oracle/_build/liboracle.so,
the plain-C restatement of the reference's hot path (ms_oracle.c, gf_oracle.c,
channel_oracle.c).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs import this; the product (channelcoding_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

V_MS, V_NMS, V_OMS, V_SCMS1, V_SCMS2, V_NMS2D, V_SPA = range(7)
VARIANTS = {"MS": 0, "NMS": 1, "OMS": 2, "SCMS1": 3, "SCMS2": 4, "2DNMS": 5, "SPA": 6, "MS_Q": 7, "NMS_Q": 8, "OMS_Q": 9}
STOP_REF_ZERO_OVERLAP, STOP_GF2_PARITY, STOP_NONE = 0, 1, 2
FAM_BCH, FAM_RS = 0, 1

_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


def build(force=False):
    """compile the C restatement (gcc, seconds)."""
    srcs = [os.path.join(_HERE, f) for f in ("ms_oracle.c", "gf_oracle.c", "channel_oracle.c")]
    if not force and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs):
        return _SO
    subprocess.check_call(["make", "-C", _HERE, "_build/liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.oracle_min_sum_batch.argtypes = [_u8p, C.c_uint, C.c_uint, _f32p, C.c_uint64, C.c_int, C.c_double,
                                           C.c_double, C.c_uint, C.c_int, _u8p, C.c_void_p, _u32p, _u8p]
        L.oracle_min_sum_fixed_batch.argtypes = [_u8p, C.c_uint, C.c_uint, _f32p, C.c_uint64, C.c_int, C.c_double,
                                                 C.c_double, C.c_uint, C.c_int, C.c_float, C.c_int, C.c_int, _u8p,
                                                 C.c_void_p, _u32p, _u8p]
        L.oracle_spa_f64_batch.argtypes = [_u8p, C.c_uint, C.c_uint, _f32p, C.c_uint64, C.c_uint, C.c_int, _u8p, C.c_void_p,
                                           _u32p, _u8p]
        L.oracle_quantise.argtypes = [C.c_float, C.c_float, C.c_int]
        L.oracle_gf_tables.argtypes = [C.c_uint, C.c_uint, _u16p, _u16p]
        L.oracle_code_new.restype = C.c_void_p
        L.oracle_code_new.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_uint, C.c_uint]
        L.oracle_code_free.argtypes = [C.c_void_p]
        L.oracle_code_params.argtypes = [C.c_void_p, C.POINTER(C.c_uint), C.POINTER(C.c_double)]
        L.oracle_code_poly.argtypes = [C.c_void_p, C.c_int, _u32p]
        L.oracle_code_H.argtypes = [C.c_void_p, C.c_uint, _u8p]
        L.oracle_encode.argtypes = [C.c_void_p, _u8p, _u8p]
        L.oracle_syndromes.argtypes = [C.c_void_p, _u8p, _u16p]
        L.oracle_hard_correct_batch.argtypes = [C.c_void_p, _u8p, C.c_uint64, _u32p, C.c_uint, _u8p, _u32p, _u8p]
        L.oracle_philox4x32_10.argtypes = [_u32p, _u32p, _u32p]
        L.oracle_sigma.restype = C.c_double
        L.oracle_sigma.argtypes = [C.c_double, C.c_double]
        L.oracle_awgn_batch.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint, C.c_float, _f32p]
        _lib = L
    return _lib


def min_sum(H, y, variant="MS", alpha=1.0, beta=0.0, max_iter=50, stop_rule=STOP_REF_ZERO_OVERLAP):
    """-> bits (frames,n) u8, L (frames,n) f32, iter (frames,) u32 [== max_iter on failure], failed u8"""
    H = np.ascontiguousarray(H, np.uint8)
    y = np.ascontiguousarray(y, np.float32).reshape(-1, H.shape[1])
    v = VARIANTS[variant] if isinstance(variant, str) else int(variant)
    f = y.shape[0]
    bits = np.zeros((f, H.shape[1]), np.uint8)
    L = np.zeros((f, H.shape[1]), np.float32)
    it = np.zeros(f, np.uint32)
    failed = np.zeros(f, np.uint8)
    lib().oracle_min_sum_batch(H, H.shape[0], H.shape[1], y, f, v, alpha, beta, max_iter, stop_rule,
                               bits, L.ctypes.data, it, failed)
    return bits, L, it, failed


def spa_f64(H, llr, max_iter=50, stop_rule=STOP_GF2_PARITY):
    """sum-product in double precision (yardstick of the float32 SPA kernel) -> bits u8, L float64, iter u32, failed u8"""
    H = np.ascontiguousarray(H, np.uint8)
    y = np.ascontiguousarray(llr, np.float32).reshape(-1, H.shape[1])
    f = y.shape[0]
    bits = np.zeros((f, H.shape[1]), np.uint8)
    L = np.zeros((f, H.shape[1]), np.float64)
    it = np.zeros(f, np.uint32)
    failed = np.zeros(f, np.uint8)
    lib().oracle_spa_f64_batch(H, H.shape[0], H.shape[1], y, f, max_iter, stop_rule, bits, L.ctypes.data, it, failed)
    return bits, L, it, failed


def min_sum_fixed(H, y, variant="MS_Q", alpha=1.0, beta=0.0, max_iter=50, stop_rule=STOP_REF_ZERO_OVERLAP, q_scale=8.0,
                  q_y_max=31, q_msg_max=31):
    """fixed-point min-sum (unpinned extension, ms_oracle.c) -> bits u8, L int32 (integer totals), iter u32, failed u8"""
    H = np.ascontiguousarray(H, np.uint8)
    y = np.ascontiguousarray(y, np.float32).reshape(-1, H.shape[1])
    v = VARIANTS[variant] if isinstance(variant, str) else int(variant)
    f = y.shape[0]
    bits = np.zeros((f, H.shape[1]), np.uint8)
    L = np.zeros((f, H.shape[1]), np.int32)
    it = np.zeros(f, np.uint32)
    failed = np.zeros(f, np.uint8)
    lib().oracle_min_sum_fixed_batch(H, H.shape[0], H.shape[1], y, f, v, alpha, beta, max_iter, stop_rule,
                                     float(q_scale), int(q_y_max), int(q_msg_max), bits, L.ctypes.data, it, failed)
    return bits, L, it, failed


def gf_tables(q, poly=0):
    size = 1 << q
    exp = np.zeros(2 * size, np.uint16)
    log = np.zeros(size, np.uint16)
    assert lib().oracle_gf_tables(q, poly, exp, log) == 0
    return exp, log


class Code:
    """BCH (family 0) / RS (family 1) code built by the C restatement of bch.h / rs.h / cyclic.h."""

    def __init__(self, family, q, t, mu=1, step=1):
        self._h = lib().oracle_code_new(family, q, t, mu, step)
        if not self._h:
            raise ValueError("code construction failed")
        out = (C.c_uint * 7)()
        rate = C.c_double()
        lib().oracle_code_params(self._h, out, C.byref(rate))
        self.n, self.l, self.k, self.dmin, self.t, self.deg_g, self.deg_h = [int(v) for v in out]
        self.rate = rate.value
        self.family, self.q = family, q

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_code_free(self._h)
            self._h = None

    def poly(self, which):
        out = np.zeros(1024, np.uint32)
        m = lib().oracle_code_poly(self._h, {"g": 0, "h": 1}[which], out)
        return out[:m].copy()

    def H(self, rows=None):
        rows = self.k if rows is None else rows
        out = np.zeros((rows, self.n), np.uint8)
        lib().oracle_code_H(self._h, rows, out)
        return out

    def encode(self, msgs):
        msgs = np.ascontiguousarray(msgs, np.uint8).reshape(-1, self.l)
        words = np.zeros((msgs.shape[0], self.n), np.uint8)
        for i in range(msgs.shape[0]):
            lib().oracle_encode(self._h, msgs[i], words[i])
        return words

    def syndromes(self, word):
        s = np.zeros(2 * self.t, np.uint16)
        lib().oracle_syndromes(self._h, np.ascontiguousarray(word, np.uint8), s)
        return s

    def hard_correct(self, words, erasures=()):
        words = np.ascontiguousarray(words, np.uint8).reshape(-1, self.n)
        erasures = list(erasures)
        er = np.asarray(erasures or [0], np.uint32)
        out = np.zeros(words.shape, np.uint8)
        nerr = np.zeros(words.shape[0], np.uint32)
        status = np.zeros(words.shape[0], np.uint8)
        lib().oracle_hard_correct_batch(self._h, words, words.shape[0], er, len(erasures), out, nerr, status)
        return out, nerr, status


def philox4x32_10(ctr, key):
    out = np.zeros(4, np.uint32)
    lib().oracle_philox4x32_10(np.asarray(ctr, np.uint32), np.asarray(key, np.uint32), out)
    return out


def sigma(rate, ebno_db):
    return lib().oracle_sigma(rate, ebno_db)


def awgn(seed, point, frame0, frames, n, sigma_f):
    y = np.zeros((frames, n), np.float32)
    lib().oracle_awgn_batch(seed, point, frame0, frames, n, np.float32(sigma_f), y)
    return y
