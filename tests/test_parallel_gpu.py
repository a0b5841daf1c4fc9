"""Two-GPU data parallelism through NCCL (needs 2 visible devices: `gpurun --gpus 2`): every rank
runs the tiny decoder on its contiguous shard of the batch, the gradients meet in ONE all-reduce of
the flat buffer, and the result must equal what a single GPU gets from the two shards run one
after the other and averaged (per-replica loss normalisation, gather-mean semantics of the
reference's data_parallel)."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = dict(B=4, T=9, S=11, F=3, O=4, P=5)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _grads_of_shard(dec, cap, ctx, lo, hi):
    for p in dec.parameters():
        p.grad = None
    c = cap[lo:hi]
    cx = {k: (v[:, lo:hi].contiguous() if not k.endswith('_mask') else v[lo:hi].contiguous())
          for k, v in ctx.items()}
    inp, tgt = c[:, :-1].contiguous(), c[:, 1:].contiguous()
    out, _ = dec({'roberta': inp}, cx)
    loss, _ = dec.adaptive_softmax.fused_loss(out, tgt)
    loss.backward()
    return {n: p.grad.detach().clone() for n, p in dec.named_parameters() if p.grad is not None}


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, 'transform-and-tell_b200'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from tell_b200 import config, synth
    from tell_b200.models import DynamicConvFacesObjectsDecoder
    from tell_b200.parallel import FlatGradients
    from tell_b200.testing import build_decoder
    config.set_precision('bf16x3')
    cfg = synth.CFG_TINY
    sd = synth.decoder_state_dict(cfg, seed=0, logit_gain=4.0)
    dec = build_decoder(cfg, DynamicConvFacesObjectsDecoder, sd).cuda().eval()
    cap, ctx = synth.decoder_inputs(cfg, **SHAPES, seed=1234)
    cap = cap.cuda()
    ctx = {k: v.cuda() for k, v in ctx.items()}
    per = SHAPES['B'] // world
    res = {}
    for dtype, tol in ((torch.float32, 1e-4), (torch.bfloat16, 1.5e-2)):   # fp32: atomics order only
        fg = FlatGradients(dec.parameters(), attach=False, dtype=dtype)
        _grads_of_shard(dec, cap, ctx, rank * per, (rank + 1) * per)
        fg.pack()
        fg.allreduce_mean()
        fg.unpack()
        got = {n: p.grad.detach().clone() for n, p in dec.named_parameters() if p.grad is not None}
        if rank == 0:      # the same two shards on ONE GPU, averaged
            parts = [_grads_of_shard(dec, cap, ctx, r * per, (r + 1) * per) for r in range(world)]
            worst = 0.0
            for n in got:
                want = sum(p[n] for p in parts) / world
                e = (got[n] - want).abs().max().item() / max(1e-6, want.abs().max().item())
                worst = max(worst, e)
            res[str(dtype)] = (worst, tol)
    dist.barrier()
    if rank == 0:
        q.put(res)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (gpurun --gpus 2)')
def test_two_gpu_allreduced_gradients_equal_single_gpu_mean():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    print('\nMEASURED two_gpu_gradients', res)
    for k, (worst, tol) in res.items():
        assert worst < tol, (k, worst)
