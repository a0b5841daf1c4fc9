"""tell_b200: B200-native (sm_100a) implementation of the Transform-and-Tell caption hot path.

Same module / model surface as the reference's `tell.modules` and `tell.models`; every op runs
in libtt_b200.so (hand-written CUDA behind a C ABI).  No CPU fallback exists.
"""
from . import config  # noqa: F401

__all__ = ['config']
