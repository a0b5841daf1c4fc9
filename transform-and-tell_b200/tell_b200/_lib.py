"""ctypes binding of libtt_b200.so (the C ABI declared in include/tt_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails the
error is raised.  The product path never routes through oracle/ or a CPU implementation.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# TT_B200_LIB: load another build of the same ABI (A/B timing of two library versions); never a fallback
LIB_PATH = os.environ.get('TT_B200_LIB') or os.path.join(_HERE, 'libtt_b200.so')

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_ll = ctypes.c_longlong
c_float = ctypes.c_float


class TtError(RuntimeError):
    pass


class TtGemmParams(ctypes.Structure):
    _fields_ = [('M', c_int), ('N', c_int), ('K', c_int),
                ('A', c_void_p), ('lda', c_ll),
                ('B', c_void_p), ('ldb', c_ll),
                ('C', c_void_p), ('ldc', c_ll),
                ('C16', c_void_p), ('ldc16', c_ll),
                ('bias', c_void_p),
                ('residual', c_void_p), ('ldr', c_ll),
                ('residual16', c_void_p), ('ldr16', c_ll),
                ('alpha', c_float), ('act', c_int), ('accumulate', c_int),
                ('m_limit', c_void_p), ('trans_a', c_int), ('trans_b', c_int), ('m_hint', c_int), ('k_limit', c_void_p),
                ('col_stats', c_void_p), ('nbatch', c_int), ('a_off0', c_int), ('a_off1', c_int),
                ('b_off0', c_int), ('b_off1', c_int), ('bias_off', c_int), ('c_off', c_ll)]


class TtAttnCtx(ctypes.Structure):
    _fields_ = [('q', c_void_p), ('k', c_void_p), ('v', c_void_p), ('bias_k', c_void_p),
                ('bias_v', c_void_p), ('mask', c_void_p), ('out', c_void_p), ('lse', c_void_p),
                ('S', c_int), ('ldq', c_ll), ('ldkv', c_ll), ('ldo', c_ll),
                ('seed', ctypes.c_ulonglong), ('dout', c_void_p), ('dq', c_void_p), ('dk', c_void_p),
                ('dv', c_void_p), ('dbias_k', c_void_p), ('dbias_v', c_void_p), ('kv_len', c_void_p),
                ('out16', c_void_p), ('dq16', c_void_p), ('ldo16', c_ll), ('ldq16', c_ll),
                ('dsum', c_void_p)]


_P4 = c_void_p * 4
_U4 = ctypes.c_ulonglong * 4


class TtLnFwdMulti(ctypes.Structure):
    _fields_ = [('h', _P4), ('gamma', _P4), ('beta', _P4), ('mean', _P4), ('rstd', _P4), ('seed', _U4),
                ('res', c_void_p), ('y', c_void_p), ('ldy', c_ll), ('y16', c_void_p), ('ldy16', c_ll),
                ('n', c_int), ('N', c_int), ('E', c_int), ('eps', ctypes.c_float),
                ('p_drop', ctypes.c_float)]


class TtLnBwdMulti(ctypes.Structure):
    _fields_ = [('x', _P4), ('mean', _P4), ('rstd', _P4), ('gamma', _P4), ('dgamma', _P4),
                ('dbeta', _P4), ('dh', _P4), ('seed', _U4), ('dy', c_void_p), ('lddy', c_ll),
                ('dx', c_void_p), ('dh16', c_void_p), ('lddh16', c_ll), ('n', c_int), ('N', c_int),
                ('E', c_int), ('p_drop', ctypes.c_float)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TtError(
                'libtt_b200.so not built (%s). Run `python -c "import __graft_entry__ as g; '
                'g.build()"` or transform-and-tell_b200/csrc/build.py. There is no CPU fallback.'
                % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.tt_last_error.restype = ctypes.c_char_p
        _lib.tt_launch_count.restype = c_ll
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().tt_last_error().decode('utf-8', 'replace')
        raise TtError('%s failed (%d): %s' % (what, rc, msg))


# When PROFILE is a list, every C-ABI call is bracketed by CUDA events on the launching stream
# (bench.py's live per-kernel timing); GEMM_FLOPS collects the algorithmic FLOPs of GEMM calls.
PROFILE = None
GEMM_FLOPS = []


def call(name, *args):
    fn = getattr(lib(), name)
    if PROFILE is None:
        check(fn(*args), name)
        return
    import torch
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    check(fn(*args), name)
    e.record()
    PROFILE.append((name, s, e))


def launch_count():
    return int(lib().tt_launch_count())


def reset_launch_count():
    lib().tt_reset_launch_count()
