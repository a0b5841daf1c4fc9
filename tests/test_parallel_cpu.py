"""world_size-2 gloo test (CPU) of the data-parallel plumbing: batch sharding + the single flat
gradient all-reduce."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from tell_b200.parallel import FlatGradients, shard_batch
    torch.manual_seed(0)
    w = torch.nn.Parameter(torch.randn(5, 3))
    tied = w                                       # tied parameter listed twice
    b = torch.nn.Parameter(torch.randn(5))
    frozen = torch.nn.Parameter(torch.randn(2), requires_grad=False)
    fg = FlatGradients([w, b, tied, frozen])
    assert fg.flat.numel() == 20
    batch = {'x': torch.arange(12.).view(4, 3), 'meta': ['a', 'b', 'c', 'd'],
             'nested': {'ids': torch.arange(4)}}
    sh = shard_batch(batch, rank, world)
    assert sh['x'].shape[0] == 2 and sh['meta'] == ['a', 'b', 'c', 'd'][2 * rank:2 * rank + 2]
    assert torch.equal(sh['nested']['ids'], torch.arange(4)[2 * rank:2 * rank + 2])
    fg.zero()
    loss = (sh['x'] @ w.t() + b).sum() * (rank + 1)
    loss.backward()
    assert w.grad.data_ptr() == fg.flat.data_ptr()          # grads accumulate into the flat views
    local = fg.flat.clone()
    fg.allreduce_mean()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    assert torch.allclose(fg.flat, sum(gathered) / world)
    # detached mode: backward writes fresh grads, pack() gathers them
    fg2 = FlatGradients([w, b, tied, frozen], attach=False)
    fg2.release()
    ((sh['x'] @ w.t() + b).sum() * (rank + 1)).backward()
    assert w.grad.data_ptr() != fg2.flat.data_ptr()
    fg2.pack()
    assert torch.allclose(fg2.flat, local)
    # bf16 payload: pack() converts while it gathers; unpack() hands the mean back as fp32 .grad
    fg3 = FlatGradients([w, b, tied, frozen], attach=False, dtype=torch.bfloat16)
    assert fg3.flat.dtype == torch.bfloat16 and fg3.flat.numel() == 20
    fg3.pack()
    assert torch.allclose(fg3.flat.float(), local, rtol=1e-2, atol=1e-2)
    fg3.allreduce_mean()
    fg3.unpack()
    assert w.grad.dtype == torch.float32
    want = (sum(gathered) / world)[:15].view(5, 3)
    assert torch.allclose(w.grad, want, rtol=2e-2, atol=2e-2)
    if rank == 0:
        out.put(fg.flat.clone())
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res.numel() == 20 and torch.isfinite(res).all()
