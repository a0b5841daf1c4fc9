"""Process-wide knobs of the B200 path.

precision
    'bf16'    bf16 tensor-core operands, fp32 accumulation (throughput mode; the reference
              trained in apex-O2 fp16, tell/training/callback_apex_trainer.py:121-125).
    'bf16x3'  every GEMM operand is split x = hi + lo (two bf16) and the product is evaluated as
              hi.hi + lo.hi + hi.lo on the same tcgen05 kernel with K tripled: ~2^-16 relative
              error, used for the "fp32 logits within 1e-3 / token-exact greedy" parity gate.
"""
import itertools
import os

import torch

precision = 'bf16'
# weight gradients on a second stream (functional.wgrad); opt-in, throughput mode only
# projected keys|values of the cross-attentions (and dL/dk, dL/dv) as bf16 in HBM (throughput mode)
kv_bf16 = True
# RoBERTa self-attention on tcgen05 tensor cores (flash_tc5.cu) instead of the mma.sync kernel
flash_tc5 = os.environ.get('TT_FLASH_TC5', '1') == '1'
# run the ResNet encoder as a parallel stream branch beside RoBERTa in Model.encode()
encoder_overlap = True
# the cross-attentions of one decoder layer (up to 4 contexts) as ONE launch per kernel type
attn_multi = os.environ.get('TT_ATTN_MULTI', '1') == '1'
# cross-attention over long contexts skips the key tiles of trailing padding (per-sample valid key count)
attn_skip_padding = os.environ.get('TT_ATTN_SKIP_PADDING', '1') == '1'
# same-shape GEMMs of one layer (the four out-projections, their dX and dW) as ONE batched launch
gemm_batched = os.environ.get('TT_GEMM_BATCHED', '1') == '1'
# producers write the bf16 operand of the consuming GEMM beside their fp32 result (twin.py, csrc/twin.cu):
# no standalone cast launches between row kernels and GEMMs; the context LayerNorms of a layer as one launch
twins = os.environ.get('TT_TWINS', '1') == '1'
wgrad_stream = 0        # 0 off, 1 bank dL/dw only (deferred join), 2 + function-local forks

# NVTX ranges around the phases of Model.forward / generate (encoders, decoder, loss, decode steps):
# visible in Nsight Systems / ncu --nvtx; off by default (a push/pop pair per phase costs ~1 us of host time)
nvtx = os.environ.get('TT_NVTX', '0') == '1'


class nvtx_range:
    """`with config.nvtx_range('name'):` -- a torch.cuda.nvtx range when config.nvtx is on."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        self.on = nvtx
        if self.on:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if self.on:
            torch.cuda.nvtx.range_pop()
        return False


# SM cap of the frozen encoders' large GEMMs inside Model.encode (0 = none).  A trainer that runs the
# encoders of batch i+1 beside the decoder forward/backward of batch i (bench.py) sets it (88 of 148
# measured best on B200): the persistent RoBERTa GEMMs stop holding the whole machine, the decoder's
# and the ResNet's latency-bound chains keep flowing, and the overlapped step gets ~5 % faster while
# the encoder graph alone is no slower.
encoder_sm_cap = int(os.environ.get('TT_ENCODER_SM_CAP', '0'))


def set_gemm_occupancy_weight(w):
    """Tile-choice objective of the GEMM dispatcher (tt_gemm_set_occupancy_weight): 0 = shortest launch
    (a chain running alone, e.g. greedy decoding), 1 = fewest SMs x time (several streams sharing the
    machine, e.g. the pipelined train step of bench.py)."""
    from . import _lib
    import ctypes
    _lib.lib().tt_gemm_set_occupancy_weight(ctypes.c_float(float(w)))


class pdl:
    """`with config.pdl():` -- kernels launched (or captured) inside use programmatic dependent launch
    (tt_set_pdl): the next kernel is scheduled while the current one still runs.  For single-chain graphs
    (the decoder forward/backward, the decode step); a graph with parallel branches gets slower."""

    def __init__(self, on=True):
        self.on = bool(on)

    def __enter__(self):
        from . import _lib
        self.prev = _lib.lib().tt_set_pdl(1 if self.on else 0)
        return self

    def __exit__(self, *exc):
        from . import _lib
        _lib.lib().tt_set_pdl(self.prev)
        return False


class gemm_sm_cap:
    """`with config.gemm_sm_cap(64):` -- large GEMMs launched (or captured) inside use at most that many
    SMs (tt_gemm_set_sm_cap), so that kernels of concurrent streams are not queued behind persistent
    GEMMs that hold the whole machine."""

    def __init__(self, sms):
        self.sms = int(sms)

    def __enter__(self):
        from . import _lib
        _lib.lib().tt_gemm_set_sm_cap(self.sms)

    def __exit__(self, *exc):
        from . import _lib
        _lib.lib().tt_gemm_set_sm_cap(0)
        return False


_seed_base = 0x5EED
_counter = itertools.count(1)
_step = None  # device-resident step counter mixed into every dropout seed


def set_precision(p):
    global precision
    assert p in ('bf16', 'bf16x3')
    precision = p


def enable_wgrad_stream(level=2):
    global wgrad_stream
    wgrad_stream = int(level)


def manual_seed(seed):
    """Re-seeds the dropout streams (the analogue of torch.manual_seed for our kernels)."""
    global _seed_base, _counter
    _seed_base = int(seed)
    _counter = itertools.count(1)


def next_seed():
    """A fresh 64-bit stream id per dropout site per call."""
    return (_seed_base * 0x9E3779B97F4A7C15 + next(_counter) * 0xBF58476D1CE4E5B9) & (2 ** 64 - 1)


def enable_device_step(device='cuda'):
    """Registers a device-side step counter with the library so that CUDA-graph replays draw
    fresh dropout masks (see tt_set_rng_step_ptr).  Returns the counter tensor."""
    global _step
    from . import _lib
    if _step is None:
        _step = torch.zeros(1, dtype=torch.int64, device=device)
        _lib.lib().tt_set_rng_step_ptr(_lib.c_void_p(_step.data_ptr()))
    return _step


def advance_device_step():
    from . import _lib
    if _step is not None:
        _lib.call('tt_rng_step_advance', _lib.c_void_p(_step.data_ptr()),
                  _lib.c_void_p(torch.cuda.current_stream().cuda_stream))


# --------------------------------------------------------------------------- zero arena (opt-in)
class ZeroArena:
    """Bump allocator over one pre-zeroed fp32 buffer for the many small accumulate-into gradients
    of a backward pass (bias column sums, LayerNorm dgamma/dbeta, bias_k/bias_v): ONE memset per
    step instead of one per tensor (~130 fill launches in the cfg-2 backward).

    Contract (why it is opt-in): slices handed out during step i become parameter .grad tensors and
    are zeroed again by reset() at the start of step i+1, so the trainer must consume (or copy)
    gradients before the next forward and must not accumulate .grad across backward calls."""

    def __init__(self, device, capacity=2 * 1024 * 1024):
        self.buf = torch.zeros(capacity, dtype=torch.float32, device=device)
        self.capacity = capacity
        self.off = 0

    def reset(self):
        self.buf.zero_()
        self.off = 0

    def take(self, n):
        n_al = (n + 3) // 4 * 4           # keep every slice 16-byte aligned
        if self.off + n_al > self.capacity:
            return None
        v = self.buf[self.off:self.off + n]
        self.off += n_al
        return v


zero_arena = None


def enable_zero_arena(device='cuda', capacity=2 * 1024 * 1024):
    global zero_arena
    if zero_arena is None:
        zero_arena = ZeroArena(device, capacity)
    return zero_arena


def disable_zero_arena():
    global zero_arena
    zero_arena = None
