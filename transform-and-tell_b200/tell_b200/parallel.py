"""Data-parallel plumbing (one process per GPU, torch.distributed).

The reference's multi-GPU path is single-process DataParallel (callback_apex_trainer.py:170-189:
replicate -> scatter batch -> gather losses, gradients reduced to device 0).  Here every rank owns
a full replica and a contiguous shard of the batch; the only collective on the data path is ONE
all-reduce (mean) of the trainable gradients, which live as views of a single flat buffer so the
collective needs no packing copies.  Frozen encoders never communicate."""
import torch
import torch.distributed as dist


def shard_batch(batch, rank, world):
    """Contiguous split of every tensor's batch dimension (dim 0); lists are sliced alike."""
    def cut(x):
        if isinstance(x, dict):
            return {k: cut(v) for k, v in x.items()}
        if torch.is_tensor(x) or isinstance(x, list):
            n = len(x)
            per = (n + world - 1) // world
            return x[rank * per:min(n, (rank + 1) * per)]
        return x
    return {k: cut(v) for k, v in batch.items()}


class FlatGradients:
    """One flat buffer holding every trainable gradient, so the step's collective is a single
    all-reduce.

    attach=True  every parameter's .grad is a view into the (fp32) buffer (backward accumulates in
                 place; call zero() first).
    attach=False gradients are produced by backward as usual (no accumulate kernels) and pack()
                 gathers them into the buffer with one batched copy before the all-reduce.
    dtype        torch.float32, or torch.bfloat16 (attach=False only): the payload of the collective
                 is half the size and pack() converts while it copies (one pass: 4 B read + 2 B
                 written per element, no separate cast).  unpack() hands the reduced values back to
                 the parameters' fp32 .grad tensors for an optimizer that wants them there."""

    def __init__(self, params, attach=True, dtype=torch.float32):
        assert dtype in (torch.float32, torch.bfloat16)
        assert not (attach and dtype != torch.float32), "attached gradients are fp32 views"
        seen, self.params = set(), []
        for p in params:
            if p.requires_grad and id(p) not in seen:      # tied weights appear once
                seen.add(id(p))
                self.params.append(p)
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=dtype, device=dev)
        self.views = []
        off = 0
        for p in self.params:
            v = self.flat[off:off + p.numel()].view_as(p)
            self.views.append(v)
            if attach:
                p.grad = v
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def pack(self):
        """flat <- concatenation of the parameters' current .grad tensors (one foreach copy)."""
        srcs = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        torch._foreach_copy_(self.views, srcs)
        return self.flat

    def unpack(self):
        """parameters' .grad <- the (reduced) flat buffer, converting back to fp32 if needed."""
        dsts, srcs = [], []
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                p.grad = torch.empty_like(p)
            if p.grad.data_ptr() != v.data_ptr():
                dsts.append(p.grad)
                srcs.append(v)
        if dsts:
            torch._foreach_copy_(dsts, srcs)

    def release(self):
        """Drop .grad tensors so the next backward writes fresh ones instead of accumulating."""
        for p in self.params:
            p.grad = None

    def allreduce_mean(self, group=None):
        """The single collective of the step: sum over ranks / world (gather-mean semantics of
        allennlp.training.util.data_parallel)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            if dist.get_backend(group) == 'nccl':
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)   # mean inside NCCL
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
                self.flat.div_(dist.get_world_size(group))
        return self.flat
