"""Spatial SM partitions (CUDA green contexts) for the two pipeline stages of a train step.

The frozen encoders of step i+1 and the decoder forward/backward of step i are independent chains
(bench.py runs them on two streams).  On one shared pool of SMs they interleave kernel by kernel: a
large encoder grid holds the block scheduler until its last wave has been issued, so the decoder's
short, latency-bound kernels queue behind it.  With the SMs split into two disjoint green contexts
each chain owns its partition and neither ever waits for the other's blocks.

`sm_partition(device_index, first_count)` -> [(n_sms, make_stream), (n_sms, make_stream)]: the
first partition holds at least `first_count` SMs (rounded by the driver to its granularity), the
second one the remaining SMs; make_stream() returns a torch stream bound to that partition.
Work captured on such a stream keeps its partition when the CUDA graph is replayed.
"""
import torch

_KEEP = []      # green contexts / raw streams stay alive for the life of the process


def _ck(res, what):
    err = res[0]
    if int(err) != 0:
        raise RuntimeError('%s failed: %s' % (what, err))
    return res[1] if len(res) == 2 else res[1:]


def sm_partition(device_index, first_count):
    from cuda.bindings import driver as drv
    torch.cuda.init()
    torch.zeros(1, device='cuda:%d' % device_index)          # primary context up
    dev = _ck(drv.cuDeviceGet(device_index), 'cuDeviceGet')
    res = _ck(drv.cuDeviceGetDevResource(dev, drv.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM),
              'cuDeviceGetDevResource')
    groups, nb, remaining = _ck(drv.cuDevSmResourceSplitByCount(1, res, 0, int(first_count)),
                                'cuDevSmResourceSplitByCount')
    parts = []
    for r in (groups[0], remaining):
        desc = _ck(drv.cuDevResourceGenerateDesc([r], 1), 'cuDevResourceGenerateDesc')
        gctx = _ck(drv.cuGreenCtxCreate(desc, dev, drv.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM),
                   'cuGreenCtxCreate')
        _KEEP.append(gctx)
        n = int(r.sm.smCount)

        def make_stream(gctx=gctx):
            s = _ck(drv.cuGreenCtxStreamCreate(gctx, drv.CUstream_flags.CU_STREAM_NON_BLOCKING, 0),
                    'cuGreenCtxStreamCreate')
            _KEEP.append(s)
            return torch.cuda.ExternalStream(int(s), device='cuda:%d' % device_index)
        parts.append((n, make_stream))
    return parts
