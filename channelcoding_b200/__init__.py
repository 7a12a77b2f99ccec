"""channelcoding_b200 -- B200-native engine for the hot path of hannesweisbach/channelcoding:
iterative min-sum-family decoding of binary BCH codes inside an AWGN Monte-Carlo sweep and batched
GF(2^q) algebraic BCH/RS decoding.  All compute happens in hand-written sm_100a CUDA kernels behind
the C ABI of include/ccgpu.h (libccgpu.so, built by `python -m channelcoding_b200.build`); there is
no CPU fallback."""
from .engine import (CAP_DMIN, CAP_ERRORS, STOP_GF2_PARITY, STOP_NONE, STOP_REF_ZERO_OVERLAP, VARIANTS,  # noqa: F401
                     Code, Context, Group, GroupCode, gf_tables, host_bch, host_from_dense, host_rs, sigma)
from ._lib import CcgpuError  # noqa: F401
from . import _lib  # noqa: F401
