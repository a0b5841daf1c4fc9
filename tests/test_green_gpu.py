"""tell_b200/green.py: two disjoint SM partitions (CUDA green contexts), streams bound to them, kernels
of the library launched (and captured into a CUDA graph) on such a stream give the same bits as on an
ordinary stream.  The bench's --green-dec-sms experiment rests on exactly this."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_sm_partition_streams_run_the_library_kernels():
    pytest.importorskip('cuda.bindings.driver')
    from tell_b200 import green, ops
    try:
        (n_a, mk_a), (n_b, mk_b) = green.sm_partition(0, 48)
    except RuntimeError as e:            # driver without green-context support
        pytest.skip(str(e))
    total = torch.cuda.get_device_properties(0).multi_processor_count
    assert n_a >= 48 and n_b >= 8 and n_a + n_b <= total, (n_a, n_b, total)
    torch.manual_seed(0)
    a = torch.randn(800, 1024, device='cuda').bfloat16()
    w = torch.randn(1024, 1024, device='cuda').bfloat16()
    ref = ops.gemm_tn(a, w)
    torch.cuda.synchronize()
    outs = []
    for mk in (mk_a, mk_b):
        s = mk()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            c = ops.gemm_tn(a, w)
        torch.cuda.current_stream().wait_stream(s)
        outs.append(c)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], ref) and torch.equal(outs[1], ref)
    # captured on a partition stream, replayed: same bits
    s = mk_b()
    out = torch.empty_like(ref)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        ops.gemm_tn(a, w, out=out)
    out.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
