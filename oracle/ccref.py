"""oracle/ccref.py -- TEST INFRASTRUCTURE.  ctypes binding of oracle/_ref/libccref*.so.

libccref.so is the reference itself (hannesweisbach/channelcoding) compiled by
oracle/build_ref.sh behind the extern "C" shim oracle/ref_shim.cc (REF-FIXED flavour: the
one-token matrix.h:50 end() fix; libccref_head.so is the as-shipped flavour).  It is used to
pin the C restatement in oracle/ (tests), to generate tests/golden/ fixtures
(oracle/make_golden.py) and as the CPU baseline of bench.py (`cpu_baseline.kind = "reference"`).
The product (channelcoding_b200/) never imports this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

FAM_BCH, FAM_RS = 0, 1
CAP_ERRORS, CAP_DMIN = 0, 1
ALG_PGZ, ALG_BM, ALG_EUKLID, ALG_SOFT0 = 0, 1, 2, 16
# soft variant ids of ref_shim.cc (soft_tag<V>)
V_MS, V_NMS, V_OMS, V_SCMS1, V_SCMS2, V_2DNMS = 0, 1, 2, 3, 4, 5
V_NMS_0915, V_OMS_0032, V_2DNMS_TUNED, V_MS_IT1, V_MS_IT5, V_NMS_IT7 = 6, 7, 8, 9, 10, 11
# (variant name, alpha, beta, max_iter) each id stands for (soft_decision.h:20-73)
VARIANT_PARAMS = {
    0: ("MS", 1.0, 0.0, 50), 1: ("NMS", 0.8, 0.0, 50), 2: ("OMS", 1.0, 0.01, 50),
    3: ("SCMS1", 1.0, 0.0, 50), 4: ("SCMS2", 1.0, 0.0, 50), 5: ("2DNMS", 1.0, 1.0, 50),
    6: ("NMS", 0.915, 0.0, 50), 7: ("OMS", 1.0, 0.032, 50), 8: ("2DNMS", 0.968, 907.0 / 125.0, 50),  # sic: beta = Beta::num / Alpha::den on *reduced* ratios
   
    9: ("MS", 1.0, 0.0, 1), 10: ("MS", 1.0, 0.0, 5), 11: ("NMS", 0.8, 0.0, 7),
}

_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")


def available(head=False):
    return os.path.exists(os.path.join(_HERE, "_ref", "libccref_head.so" if head else "libccref.so"))


class Ref:
    def __init__(self, head=False):
        path = os.path.join(_HERE, "_ref", "libccref_head.so" if head else "libccref.so")
        self.lib = lib = C.CDLL(path)
        lib.ccref_code_params.argtypes = [C.c_int] * 4 + [C.POINTER(C.c_uint), C.POINTER(C.c_double)]
        lib.ccref_code_H.argtypes = [C.c_int] * 5 + [_u8p]
        lib.ccref_code_poly.argtypes = [C.c_int] * 5 + [_u32p, C.c_int]
        lib.ccref_code_to_string.argtypes = [C.c_int] * 5 + [C.c_char_p, C.c_int]
        lib.ccref_gf_tables.argtypes = [C.c_int, _u16p, _u16p]
        lib.ccref_min_sum.argtypes = [C.c_int, _u8p, C.c_uint, C.c_uint, _f32p, C.c_uint64,
                                      _u8p, _f32p, _u32p, _u8p]
        lib.ccref_soft_correct.argtypes = [C.c_int] * 5 + [_f32p, C.c_uint64, _u8p, _u8p]
        lib.ccref_hard_correct.argtypes = [C.c_int] * 5 + [_u8p, C.c_uint64, _u32p, C.c_uint32, _u8p, _u8p]
        lib.ccref_encode.argtypes = [C.c_int] * 4 + [_u8p, C.c_uint64, _u8p]
        lib.ccref_awgn_baseline.argtypes = [C.c_int] * 5 + [C.c_double, C.c_uint64, C.c_double, C.c_int,
                                                           C.c_uint64, C.POINTER(C.c_uint64),
                                                           C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
        self.flavour = "REF-FIXED" if lib.ccref_flavour() == 1 else "REF-HEAD"

    def params(self, fam, q, kind, value):
        out = (C.c_uint * 7)()
        rate = C.c_double()
        if self.lib.ccref_code_params(fam, q, kind, value, out, C.byref(rate)) != 0:
            raise KeyError((fam, q, kind, value))
        keys = ["n", "l", "k", "dmin", "t", "H_rows", "H_alt_rows"]
        d = dict(zip(keys, [int(v) for v in out]))
        d["rate"] = rate.value
        return d

    def H(self, fam, q, kind, value, alt=False):
        p = self.params(fam, q, kind, value)
        rows = p["H_alt_rows"] if alt else p["H_rows"]
        out = np.zeros((rows, p["n"]), np.uint8)
        assert self.lib.ccref_code_H(fam, q, kind, value, int(alt), out) == 0
        return out

    def poly(self, fam, q, kind, value, which):
        out = np.zeros(1024, np.uint32)
        m = self.lib.ccref_code_poly(fam, q, kind, value, {"g": 0, "h": 1}[which], out, 1024)
        assert m >= 0
        return out[:m].copy()

    def to_string(self, fam, q, kind, value, alg):
        buf = C.create_string_buffer(128)
        if self.lib.ccref_code_to_string(fam, q, kind, value, alg, buf, 128) != 0:
            raise KeyError((fam, q, kind, value, alg))
        return buf.value.decode()

    def gf_tables(self, q):
        size = 1 << q
        exp = np.zeros(2 * size, np.uint16)
        log = np.zeros(size, np.uint16)
        assert self.lib.ccref_gf_tables(q, exp, log) == 0
        return exp, log

    def min_sum(self, variant, H, y):
        H = np.ascontiguousarray(H, np.uint8)
        y = np.ascontiguousarray(y, np.float32).reshape(-1, H.shape[1])
        frames = y.shape[0]
        bits = np.zeros((frames, H.shape[1]), np.uint8)
        L = np.zeros((frames, H.shape[1]), np.float32)
        it = np.zeros(frames, np.uint32)
        failed = np.zeros(frames, np.uint8)
        assert self.lib.ccref_min_sum(variant, H, H.shape[0], H.shape[1], y, frames, bits, L, it, failed) == 0
        return bits, L, it, failed

    def soft_correct(self, fam, q, kind, value, variant, y):
        n = self.params(fam, q, kind, value)["n"]
        y = np.ascontiguousarray(y, np.float32).reshape(-1, n)
        bits = np.zeros(y.shape, np.uint8)
        failed = np.zeros(y.shape[0], np.uint8)
        assert self.lib.ccref_soft_correct(fam, q, kind, value, variant, y, y.shape[0], bits, failed) == 0
        return bits, failed

    def hard_correct(self, fam, q, kind, value, alg, words, erasures=()):
        n = self.params(fam, q, kind, value)["n"]
        words = np.ascontiguousarray(words, np.uint8).reshape(-1, n)
        er = np.asarray(list(erasures), np.uint32)
        if er.size == 0:
            er = np.zeros(1, np.uint32)
            ner = 0
        else:
            ner = er.size
        out = np.zeros(words.shape, np.uint8)
        status = np.zeros(words.shape[0], np.uint8)
        assert self.lib.ccref_hard_correct(fam, q, kind, value, alg, words, words.shape[0], er, ner, out, status) == 0
        return out, status

    def encode(self, fam, q, kind, value, msgs):
        p = self.params(fam, q, kind, value)
        msgs = np.ascontiguousarray(msgs, np.uint8).reshape(-1, p["l"])
        words = np.zeros((msgs.shape[0], p["n"]), np.uint8)
        assert self.lib.ccref_encode(fam, q, kind, value, msgs, msgs.shape[0], words) == 0
        return words

    def awgn_baseline(self, fam, q, kind, value, alg, ebno_db, seed=0, seconds=10.0, threads=1,
                      max_frames_per_thread=0):
        frames, werr, el = C.c_uint64(), C.c_uint64(), C.c_double()
        rc = self.lib.ccref_awgn_baseline(fam, q, kind, value, alg, ebno_db, seed, seconds, threads,
                                          max_frames_per_thread, C.byref(frames), C.byref(werr), C.byref(el))
        assert rc == 0, rc
        return frames.value, werr.value, el.value
