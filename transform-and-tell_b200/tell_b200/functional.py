"""torch.autograd glue around the C-ABI kernels.

Each Function's forward AND backward call only libtt_b200.so kernels (through tell_b200.ops);
torch is used for tensor allocation, views and autograd bookkeeping.  The backward formulas are
the ones the reference leaves to torch autograd.
"""
import torch
from torch.autograd import Function

from . import config, ops, twin, weight_bank


def operand(x, side, transpose=False):
    """fp32 [R,C] -> bf16 tcgen05 GEMM operand (K-contiguous).  side: 'a' (activations/left) or
    'b' (weights/right); in bf16x3 mode the two sides get complementary hi/lo patterns."""
    if config.precision == 'bf16':
        split = 0
        if side == 'a' and not transpose and config.twins:
            t16 = twin.get(x)                       # written by the kernel that produced x
            if t16 is not None:
                return t16
        if side == 'b' and not transpose and weight_bank.ACTIVE is not None:
            hit = weight_bank.ACTIVE.single(x)      # prepared by the step's one weight_prep launch
            if hit is not None:
                return hit
    else:
        split = 1 if side == 'a' else 2
    return ops.cast_bf16(x, transpose=transpose, split=split)


def _c(x):
    return x if x.is_contiguous() else x.contiguous()


# Gradients of tied matrices produced earlier in the running backward (AdaptiveLossFn: the adaptive
# softmax shares its word matrices with the adaptive input embedding, tie_adaptive_weights).  A later
# backward node that contributes to the same parameter (EmbedFn) accumulates INTO that tensor and
# returns None instead of materialising a second dense gradient that autograd would then have to
# zero-fill and add (206 MB of fills + 600 MB of add traffic per step for the 50265 x 1024 tables).
_TIED_GRADS = {}


class _WGradState:
    stream = None      # the weight-gradient stream of this process
    main = None        # stream the running backward was forked from
    pending = False    # work has been forked since the last join
    keep = []          # inputs of forked kernels, kept alive until the join


class wgrad:
    """`with wgrad(*inputs):` -- kernels launched inside run on the weight-gradient stream.

    In the backward of y = x.W^T only dL/dx is on the critical path; dL/dW and the bias column sums
    feed nothing before the end of the backward pass.  They are forked onto a second stream (a
    parallel branch when the step is captured in a CUDA graph) so that the small, latency-bound
    decoder kernels of the two chains overlap.  The branch is joined (a) before the batched
    weight-norm backward reads the bank's dL/dw buffer and (b) by an autograd final callback, i.e.
    before loss.backward() returns -- callers see ordinary stream semantics.
    `inputs` are the tensors the forked kernels read; they are kept alive until the join so that
    the allocator cannot hand their memory to later main-stream kernels.
    Opt-in (config.wgrad_stream) and throughput mode only.  Deferred joins (level 1) are only used
    for the bank's dL/dw buffers, which autograd never touches; everything else that may be summed,
    sliced or copied by autograd on the main stream uses local=True (level 2) and joins before the
    backward function returns."""

    def __init__(self, *inputs, local=False):
        self.inputs = inputs
        self.cm = None
        self.level = 2 if local else 1

    def __enter__(self):
        if not (config.wgrad_stream >= self.level and config.precision == 'bf16'):
            return self
        st = _WGradState
        main = torch.cuda.current_stream()
        if st.stream is None or st.stream.device != main.device:
            st.stream = torch.cuda.Stream(device=main.device)
        if main == st.stream:          # nested use: already on the branch
            return self
        if not st.pending:
            from torch.autograd import Variable
            Variable._execution_engine.queue_callback(wgrad_join)   # only legal inside backward
            st.pending, st.main = True, main
        st.stream.wait_stream(main)
        st.keep.extend(t for t in self.inputs if t is not None)
        self.cm = torch.cuda.stream(st.stream)
        self.cm.__enter__()
        return self

    def __exit__(self, *exc):
        if self.cm is not None:
            self.cm.__exit__(*exc)
        return False


def wgrad_join(local=False):
    """Main stream waits for everything forked with `wgrad` so far.  local=True: the join that
    closes a fork made with wgrad(local=True) inside one backward function (outputs that autograd
    may touch on the main stream right after the function returns)."""
    st = _WGradState
    if local and config.wgrad_stream < 2:
        return
    if st.pending:
        st.main.wait_stream(st.stream)
        cur = torch.cuda.current_stream()
        if cur != st.main and cur != st.stream:
            cur.wait_stream(st.stream)
        st.pending = False
        st.keep = []


def kv16_ok(head_dim, need_weights=False):
    """Projected keys|values (and their gradients) as bf16 in HBM: throughput mode with the
    tensor-core attention kernels (head_dim 64); the head-averaged attention weights of eval mode
    are computed by an fp32 kernel, so that path keeps fp32 keys."""
    return config.precision == 'bf16' and config.kv_bf16 and head_dim == 64 and not need_weights


def _as_operand(d):
    """A gradient that is already a bf16 GEMM operand (bf16 dL/dkv slab) is used as it is."""
    return d if d.dtype == torch.bfloat16 else operand(d, 'a')


def _tw():
    """Producers emit bf16 operand twins (throughput mode only: bf16x3 operands are hi/lo splits)."""
    return config.precision == 'bf16' and config.twins


def _fast():
    """bf16 throughput mode: backward GEMMs read the forward's bf16 operands in place through
    MN-major UMMA descriptors (trans_a / trans_b) -- no transposed copies, no weight re-casts.
    In bf16x3 mode the hi/lo segments run along the contraction axis, which differs between
    forward and backward, so that mode re-casts (explicit transposes)."""
    return config.precision == 'bf16'


class LinearFn(Function):
    """y = act(alpha * (x @ w^T + bias)); x [M,K], w [N,K] (F.linear / GehringLinear / in_proj)."""

    @staticmethod
    def forward(ctx, x, w, bias, alpha=1.0, act=ops.ACT_NONE, twin_out=False):
        a16 = operand(x, 'a')
        b16 = operand(w, 'b')
        if twin_out and _tw() and w.shape[0] % 8 == 0:
            # the consumer of y is another GEMM: its bf16 operand comes out of this epilogue
            y, y16 = ops.gemm_tn(a16, b16, bias=bias, alpha=alpha, act=act, want16=True)
            twin.put(y, y16)
        else:
            y = ops.gemm_tn(a16, b16, bias=bias, alpha=alpha, act=act)
        ctx.alpha, ctx.act, ctx.has_bias, ctx.fast = alpha, act, bias is not None, _fast()
        ctx.bank, ctx.wkey = weight_bank.ACTIVE, w.data_ptr()
        keep_y = y if act != ops.ACT_NONE else None
        if ctx.fast:
            ctx.save_for_backward(a16, b16, keep_y)
        else:
            ctx.save_for_backward(x, w, keep_y)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        dy = _c(dy)
        dy16 = None
        if ctx.act == ops.ACT_RELU:
            if ctx.fast and _tw() and dy.shape[1] % 8 == 0:
                dy, dy16 = ops.relu_bwd_tw(dy, y)        # the masked gradient and its operand twin
            else:
                dy = ops.relu_bwd(dy, y)
        elif ctx.act != ops.ACT_NONE:
            raise NotImplementedError('backward of fused GELU is not needed on this path')
        dx = dw = db = None
        if ctx.fast:
            if dy16 is None:
                dy16 = operand(dy, 'a')                  # [M, N]: K-major for dx, MN-major for dw
            if ctx.needs_input_grad[0]:
                dx = ops.gemm_tn(dy16, w, alpha=ctx.alpha, trans_b=True)          # dy . W
            if ctx.needs_input_grad[1]:
                # weight-normalised matrices: dL/dw lands in the bank's flat buffer, where the
                # single batched wnorm backward reads it
                dest = ctx.bank.claim_dw(ctx.wkey) if ctx.bank is not None else None
                if dest is not None:
                    # nothing reads the bank's dL/dw buffer before the batched weight-norm
                    # backward (which joins): off the critical path, on the wgrad stream
                    with wgrad(dy16, x):
                        dw = ops.gemm_tn(dy16, x, out=dest, alpha=ctx.alpha, trans_a=True,
                                         trans_b=True)                              # dy^T . x
                else:
                    dw = ops.gemm_tn(dy16, x, alpha=ctx.alpha, trans_a=True, trans_b=True)
        else:
            if ctx.needs_input_grad[0]:
                dx = ops.gemm_tn(operand(dy, 'a'), operand(w, 'b', transpose=True), alpha=ctx.alpha)
            if ctx.needs_input_grad[1]:
                dw = ops.gemm_tn(operand(dy, 'a', transpose=True), operand(x, 'b', transpose=True),
                                 alpha=ctx.alpha)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = ops.colsum(dy, scale=ctx.alpha)
        return dx, dw, db, None, None, None


class KVProjFn(Function):
    """kv [R, 2E] = x @ [Wk; Wv]^T + [bk | bv]: the key and value projections of one context as ONE
    GEMM whose output the attention kernels read with stride 2E (multi_head.py:500-518)."""

    @staticmethod
    def forward(ctx, x, wk, wv, bias_kv, kv16=False):
        a16 = operand(x, 'a')
        w16 = concat_rows_operand([wk, wv], 'b', x.device)          # [2E, kdim]
        kv = ops.gemm_tn(a16, w16, bias=bias_kv, want32=not kv16, want16=kv16)
        ctx.has_bias, ctx.fast, ctx.E = bias_kv is not None, _fast(), wk.shape[0]
        if ctx.fast:
            ctx.save_for_backward(a16, w16)
        else:
            ctx.save_for_backward(x, wk, wv)
        return kv

    @staticmethod
    def backward(ctx, dkv):
        dkv = _c(dkv)
        E = ctx.E
        dx = dW = db = None
        need_w = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        if ctx.fast:
            x, w = ctx.saved_tensors
            d16 = _as_operand(dkv)
            if ctx.needs_input_grad[0]:
                dx = ops.gemm_tn(d16, w, trans_b=True)
            if need_w:
                dW = ops.gemm_tn(d16, x, trans_a=True, trans_b=True)
        else:
            x, wk, wv = ctx.saved_tensors
            if ctx.needs_input_grad[0]:
                dx = ops.gemm_tn(operand(dkv, 'a'), concat_kT_operand([wk, wv], 'b', dkv.device))
            if need_w:
                dW = ops.gemm_tn(operand(dkv, 'a', transpose=True), operand(x, 'b', transpose=True))
        if ctx.has_bias and ctx.needs_input_grad[3]:
            db = ops.colsum(dkv)
        return dx, (dW[:E] if dW is not None else None), (dW[E:] if dW is not None else None), db, None


class GradSlab:
    """Side channel between AllLayerKVProjFn and the attention backward: the L per-layer dL/dkv
    blocks are written straight into column blocks of ONE [R, L*2E] buffer, so the projection's
    backward is one operand cast + one dW GEMM (+ one dx GEMM) over all layers, with no
    gather copies.  Allocated lazily by the first attention backward that needs it."""

    def __init__(self, rows, width, n, device, dtype=torch.float32):
        self.rows, self.width, self.n, self.device, self.dtype = rows, width, n, device, dtype
        self.buf = None
        self.written = set()

    def block(self, l):
        if self.buf is None:
            self.buf = torch.empty((self.rows, self.n * self.width), dtype=self.dtype,
                                   device=self.device)
        self.written.add(l)
        return self.buf[:, l * self.width:(l + 1) * self.width]


class AllLayerKVProjFn(Function):
    """The key|value projections of ONE context for ALL L decoder layers as one GEMM:
    KV [R, L*2E] = x @ [Wk_0; Wv_0; ...; Wk_{L-1}; Wv_{L-1}]^T + bias  (multi_head.py:500-518, once
    per layer in the reference).  The context does not depend on the layer, so x is cast once and
    the N dimension is L times wider (8192 x 8192 x 1024 for the article instead of four
    8192 x 2048 x 1024 GEMMs).  Returns L views [R, 2E] (row stride L*2E), read by stride."""

    @staticmethod
    def forward(ctx, x, L, slab, kv16, *args):
        wks, wvs, biases = args[:L], args[L:2 * L], args[2 * L:3 * L]
        E = wks[0].shape[0]
        a16 = operand(x, 'a')
        mats = []
        for l in range(L):
            mats += [wks[l], wvs[l]]
        w16 = concat_rows_operand(mats, 'b', x.device)                  # [L*2E, kdim]
        bias = torch.cat([b.reshape(-1) for b in biases]) if biases[0] is not None else None
        KV = ops.gemm_tn(a16, w16, bias=bias, want32=not kv16, want16=kv16)
        ctx.cfg = (L, E, bias is not None, _fast())
        ctx.slab = slab
        if ctx.cfg[3]:
            ctx.save_for_backward(a16, w16)
        else:
            ctx.save_for_backward(x, *mats)
        return tuple(KV[:, l * 2 * E:(l + 1) * 2 * E] for l in range(L))

    @staticmethod
    def backward(ctx, *dkvs):
        L, E, has_bias, fast = ctx.cfg
        slab = ctx.slab
        for l, d in enumerate(dkvs):
            if d is None:
                slab.block(l).zero_()
            elif slab.buf is None or d.data_ptr() != slab.block(l).data_ptr():
                slab.block(l).copy_(d)
        D = slab.buf
        slab.buf = None                       # the buffer belongs to this backward only
        slab.written = set()
        dx = dW = None
        need_w = any(ctx.needs_input_grad[4:4 + 2 * L])
        if fast:
            a16, w16 = ctx.saved_tensors
            d16 = _as_operand(D)
            if ctx.needs_input_grad[0]:
                dx = ops.gemm_tn(d16, w16, trans_b=True)
            if need_w:
                dW = ops.gemm_tn(d16, a16, trans_a=True, trans_b=True)          # [L*2E, kdim]
        else:
            x, mats = ctx.saved_tensors[0], ctx.saved_tensors[1:]
            if ctx.needs_input_grad[0]:
                dx = ops.gemm_tn(operand(D, 'a'), concat_kT_operand(list(mats), 'b', D.device))
            if need_w:
                dW = ops.gemm_tn(operand(D, 'a', transpose=True), operand(x, 'b', transpose=True))
        dwk = tuple(dW[(2 * l) * E:(2 * l + 1) * E] if dW is not None else None for l in range(L))
        dwv = tuple(dW[(2 * l + 1) * E:(2 * l + 2) * E] if dW is not None else None for l in range(L))
        dbs = (None,) * L
        if has_bias:
            db = ops.colsum(D)
            dbs = tuple(db[l * 2 * E:(l + 1) * 2 * E] for l in range(L))
        return (dx, None, None, None) + dwk + dwv + dbs


class InProjSplitFn(Function):
    """in_proj_weight [3E,E] -> (Wq, Wk, Wv) and in_proj_bias [3E] -> (bq, bk|bv) as views
    (multi_head.py:491-518 slices them per projection).  Done with plain autograd slicing every
    slice costs a zero-filled full-size gradient, a copy and an add (SliceBackward + accumulation:
    ~100 tiny launches per step for the 16 attention modules); here the backward assembles each
    parameter's gradient with ONE concatenation."""

    @staticmethod
    def forward(ctx, w, b, E):
        outs = []
        if w is not None:
            outs += [w[:E], w[E:2 * E], w[2 * E:]]
        if b is not None:
            outs += [b[:E], b[E:]]
        ctx.cfg = (w is not None, b is not None, E)
        ctx.meta = (w.shape[1] if w is not None else 0, (w if w is not None else b).device)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        has_w, has_b, E = ctx.cfg
        K, dev = ctx.meta
        dw = db = None
        i = 0
        if has_w:
            parts = [g if g is not None else torch.zeros((E, K), dtype=torch.float32, device=dev)
                     for g in grads[0:3]]
            dw = torch.cat(parts, 0)
            i = 3
        if has_b:
            sizes = (E, 2 * E)
            parts = [g.reshape(-1) if g is not None else torch.zeros(n, dtype=torch.float32, device=dev)
                     for g, n in zip(grads[i:i + 2], sizes)]
            db = torch.cat(parts, 0)
        return dw, db, None


class WeightNormFn(Function):
    """w = g * v / ||v||_row  (nn.utils.weight_norm dim=0; linear.py:30-34)."""

    @staticmethod
    def forward(ctx, v, g):
        w, norm = ops.wnorm_fwd(v, g)
        ctx.save_for_backward(v, g, norm)
        return w

    @staticmethod
    def backward(ctx, dw):
        v, g, norm = ctx.saved_tensors
        dv, dg = ops.wnorm_bwd(_c(dw), v, g, norm)
        return dv, dg


class ResidualLayerNormFn(Function):
    """y = LayerNorm(res + dropout(h)).  h (a GEMM output nobody else reads) is overwritten with
    the pre-norm sum, which is what the backward needs."""

    @staticmethod
    def forward(ctx, h, res, gamma, beta, p, seed, eps=1e-5):
        ctx.tw = _tw() and h.shape[1] % 8 == 0 and h.shape[1] <= 1024 and h.is_contiguous() and \
            (res is None or res.is_contiguous())
        if ctx.tw:
            y, y16, means, rstds = ops.ln_fwd_multi([h], res, [gamma], [beta], eps, p, [seed])
            mean, rstd = means[0], rstds[0]
            twin.put(y, y16)
        else:
            y, mean, rstd = ops.ln_fwd(h, res, gamma, beta, eps, p, seed)
        ctx.p, ctx.seed, ctx.has_res = p, seed, res is not None
        ctx.save_for_backward(h, mean, rstd, gamma)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd, gamma = ctx.saved_tensors
        dgamma, dbeta = ops.zeros_f32(gamma.shape, gamma), ops.zeros_f32(gamma.shape, gamma)
        if ctx.tw:
            drop = ctx.p > 0
            dx, dhs, dh16 = ops.ln_bwd_multi(_c(dy), [x], [mean], [rstd], [gamma], ctx.p, [ctx.seed],
                                             want_dx=ctx.has_res or not drop, want_dh32=drop,
                                             dgammas=[dgamma], dbetas=[dbeta])
            dh = dhs[0] if drop else dx
            twin.put(dh, dh16)                     # dh feeds the backward GEMMs of the branch's last Linear
            return dh, (dx if ctx.has_res else None), dgamma, dbeta, None, None, None
        if ctx.p > 0:
            dx, dh = ops.ln_bwd(_c(dy), x, mean, rstd, gamma, ctx.p, ctx.seed,
                                want_dx=ctx.has_res, dgamma=dgamma, dbeta=dbeta)
        else:
            dx, _ = ops.ln_bwd(_c(dy), x, mean, rstd, gamma, 0.0, 0, want_dh=False,
                               dgamma=dgamma, dbeta=dbeta)
            dh = dx
        return dh, (dx if ctx.has_res else None), dgamma, dbeta, None, None, None


class ContextLayerNormFn(Function):
    """The four parallel branches of decoder_faces_objects.py:272-352 after their out_proj:
    Y[:, c*E:(c+1)*E] = LN_c(X + dropout(h_c)), written straight into the concatenated buffer that
    context_fc consumes (replaces torch.cat, :354)."""

    @staticmethod
    def forward(ctx, X, p, seeds, eps, n, *args):
        hs, gammas, betas = args[:n], args[n:2 * n], args[2 * n:3 * n]
        N, E = X.shape
        Y = torch.empty((N, n * E), dtype=torch.float32, device=X.device)
        saved = [X]
        for c in range(n):
            _, mean, rstd = ops.ln_fwd(hs[c], X, gammas[c], betas[c], eps, p, seeds[c],
                                       out=Y[:, c * E:(c + 1) * E])
            saved += [hs[c], mean, rstd, gammas[c]]
        ctx.n, ctx.p, ctx.seeds = n, p, seeds
        ctx.save_for_backward(*saved)
        return Y

    @staticmethod
    def backward(ctx, dY):
        saved = ctx.saved_tensors
        X = saved[0]
        n = ctx.n
        N, E = X.shape
        dY = _c(dY)
        dX = None
        dhs, dgs, dbs = [], [], []
        for c in range(n):
            x, mean, rstd, gamma = saved[1 + 4 * c:5 + 4 * c]
            dg, db = ops.zeros_f32(gamma.shape, gamma), ops.zeros_f32(gamma.shape, gamma)
            dyc = dY[:, c * E:(c + 1) * E]
            if ctx.p > 0:
                dx, dh = ops.ln_bwd(dyc, x, mean, rstd, gamma, ctx.p, ctx.seeds[c], dgamma=dg,
                                    dbeta=db)
            else:
                dx, _ = ops.ln_bwd(dyc, x, mean, rstd, gamma, 0.0, 0, want_dh=False, dgamma=dg,
                                   dbeta=db)
                dh = dx
            if dX is None:
                dX = dx if ctx.p > 0 else dx.clone()
            else:
                ops.axpby(dx, dX, 1.0, 1.0)
            dhs.append(dh)
            dgs.append(dg)
            dbs.append(db)
        return (dX, None, None, None, None) + tuple(dhs) + tuple(dgs) + tuple(dbs)


class Transpose01Fn(Function):
    """[A,B,C] -> [B,A,C] (the decoder's T x B x C <-> B x T x C flips)."""

    @staticmethod
    def forward(ctx, x):
        return ops.transpose01(_c(x))

    @staticmethod
    def backward(ctx, dy):
        return ops.transpose01(_c(dy))


class GLUFn(Function):
    @staticmethod
    def forward(ctx, h):
        ctx.save_for_backward(h)
        ctx.tw = _tw() and h.is_contiguous() and (h.shape[1] // 2) % 8 == 0
        if ctx.tw:
            out, out16 = ops.glu_fwd_tw(h)
            twin.put(out, out16)                   # operand of the DynamicConv filter projection
            return out
        return ops.glu_fwd(h)

    @staticmethod
    def backward(ctx, dout):
        (h,) = ctx.saved_tensors
        if ctx.tw:
            dh, dh16 = ops.glu_bwd_tw(_c(dout), h)
            twin.put(dh, dh16)                     # operand of linear1's backward GEMMs
            return dh
        return ops.glu_bwd(_c(dout), h)


class DropoutFn(Function):
    """F.dropout; x [N,C].  twin_out: the consumer is a GEMM (input dropout -> linear1, relu dropout
    -> fc2), so the bf16 operand is written by the same pass."""

    @staticmethod
    def forward(ctx, x, p, seed, twin_out=False):
        ctx.p, ctx.seed = p, seed
        x = _c(x)
        if twin_out and _tw() and x.dim() == 2 and x.shape[1] % 8 == 0:
            y, y16 = ops.dropout_tw(x, p, seed)
            twin.put(y, y16)
            return y
        return ops.dropout(x, p, seed)

    @staticmethod
    def backward(ctx, dy):
        return ops.dropout(_c(dy), ctx.p, ctx.seed), None, None, None


class DynConvFn(Function):
    """x [T,B,C], z [T,B,H*K] -> out [T,B,C]  (dynamic.py:302-335)."""

    @staticmethod
    def forward(ctx, x, z, H, K, softmax, p, seed):
        ctx.tw = _tw() and x.shape[2] % 8 == 0
        if ctx.tw:
            out, probs, out16 = ops.dynconv_fwd(x, z, H, K, softmax, p, seed, twin=True)
            twin.put(out.view(-1, x.shape[2]), out16)   # operand of linear2 (which sees out as [T*B, C])
        else:
            out, probs = ops.dynconv_fwd(x, z, H, K, softmax, p, seed)
        ctx.cfg = (H, K, softmax, p, seed)
        ctx.save_for_backward(x, probs)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, probs = ctx.saved_tensors
        H, K, softmax, p, seed = ctx.cfg
        T, B, _ = x.shape
        if ctx.tw:
            dx, dz, dz16 = ops.dynconv_bwd(_c(dout), x, probs, H, K, softmax, p, seed, twin=True)
            dz2 = dz.view(T * B, H * K)
            twin.put(dz2, dz16)                    # operand of the filter projection's backward GEMMs
            return dx, dz.view(T, B, H * K), None, None, None, None, None
        dx, dz = ops.dynconv_bwd(_c(dout), x, probs, H, K, softmax, p, seed)
        return dx, dz.view(T, B, H * K), None, None, None, None, None


class LightConvFn(Function):
    """LightweightConv1dTBC (lightweight.py:88-240): static taps w [H,K] shared over (t,b)."""

    @staticmethod
    def forward(ctx, x, w, H, K, softmax, p, seed):
        out, probs = ops.dynconv_fwd(x, w, H, K, softmax, p, seed, broadcast=True)
        ctx.cfg = (H, K, softmax, p, seed)
        ctx.save_for_backward(x, probs)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, probs = ctx.saved_tensors
        H, K, softmax, p, seed = ctx.cfg
        dx, dz = ops.dynconv_bwd(_c(dout), x, probs, H, K, softmax, p, seed)
        T, B, _ = x.shape
        dw = ops.colsum(dz.view(T * B, H * K)).view(H, K)
        return dx, dw, None, None, None, None, None


class AttentionFn(Function):
    """q [T*B,E] (scaled), kv [S*B,2E] (or None when the context is empty) -> out [T*B,E]."""

    @staticmethod
    def forward(ctx, q, kv, bias_k, bias_v, mask, T, B, S, H, zero_row, p, seed, need_weights):
        E = q.shape[1]
        D = E // H
        k = kv[:, :E] if S > 0 else None
        v = kv[:, E:] if S > 0 else None
        bk = bias_k.view(-1) if bias_k is not None else None
        bv = bias_v.view(-1) if bias_v is not None else None
        tc = config.precision == 'bf16' and D == 64      # tensor-core kernels in throughput mode
        out, lse = ops.attn_fwd(q, k, v, bk, bv, mask, T, B, S, H, D, zero_row, p, seed, tc=tc)
        ctx.cfg = (T, B, S, H, D, zero_row, p, seed, tc)
        ctx.save_for_backward(q, kv, bias_k, bias_v, mask, out, lse)
        weights = None
        if need_weights:
            weights = ops.attn_avg_weights(q, k, bk, mask, lse, T, B, S, H, D, zero_row)
            ctx.mark_non_differentiable(weights)
        return out, weights

    @staticmethod
    def backward(ctx, dout, _dweights):
        q, kv, bias_k, bias_v, mask, out, lse = ctx.saved_tensors
        T, B, S, H, D, zero_row, p, seed, tc = ctx.cfg
        E = H * D
        dq = torch.empty_like(q)
        dkv = torch.empty_like(kv) if S > 0 else None
        dbk = ops.zeros_f32(bias_k.shape, bias_k) if bias_k is not None else None
        dbv = ops.zeros_f32(bias_v.shape, bias_v) if bias_v is not None else None
        ops.attn_bwd(_c(dout), q, kv[:, :E] if S > 0 else None, kv[:, E:] if S > 0 else None,
                     bias_k.view(-1) if bias_k is not None else None,
                     bias_v.view(-1) if bias_v is not None else None, mask, out, lse, dq,
                     dkv[:, :E] if S > 0 else None, dkv[:, E:] if S > 0 else None,
                     dbk, dbv, T, B, S, H, D, zero_row, p, seed, tc=tc)
        return (dq, dkv, dbk, dbv) + (None,) * 9


def concat_rows_operand(mats, side, device):
    """One bf16 operand [sum_i R_i, K(*rep)]: the row-wise concatenation of `mats` ([R_i, K])."""
    K = mats[0].shape[1]
    rep = 1 if config.precision == 'bf16' else 3
    if rep == 1 and side == 'b' and weight_bank.ACTIVE is not None:
        hit = weight_bank.ACTIVE.group('rows', mats)
        if hit is not None:
            return hit
    split = 0 if rep == 1 else (1 if side == 'a' else 2)
    buf = ops.bf16_buffer(sum(m.shape[0] for m in mats), K * rep, device)
    r0 = 0
    for m in mats:
        ops.cast_bf16(m, split=split, out=buf[r0:r0 + m.shape[0]])
        r0 += m.shape[0]
    return buf


class FusedQProjFn(Function):
    """The query projections of all n contexts of a layer share their input (decoder_faces_objects.py
    :272-352 feeds the same X to every attention): Q_all [N, n*E] = alpha * (x @ [Wq_1;..;Wq_n]^T + b)
    as ONE GEMM (multi_head.py:491-498 in_proj_q, :353 q *= scaling)."""

    @staticmethod
    def forward(ctx, x, alpha, n, *args):
        ws, bs = args[:n], args[n:2 * n]
        a16 = operand(x, 'a')
        w16 = concat_rows_operand(list(ws), 'b', x.device)
        bias = torch.cat([b.reshape(-1) for b in bs]) if bs[0] is not None else None
        y = ops.gemm_tn(a16, w16, bias=bias, alpha=alpha)
        ctx.cfg = (alpha, n, ws[0].shape[0], bias is not None, _fast())
        if ctx.cfg[4]:
            ctx.save_for_backward(a16, w16)
        else:
            ctx.save_for_backward(x, *ws)
        return y

    @staticmethod
    def backward(ctx, dy):
        alpha, n, E, has_bias, fast = ctx.cfg
        dy = _c(dy)
        if fast:
            a16, w16 = ctx.saved_tensors
            d16 = operand(dy, 'a')
            db = None
            with wgrad(d16, a16, dy, local=True):      # dW, db alongside dx; joined before return
                dW = ops.gemm_tn(d16, a16, alpha=alpha, trans_a=True, trans_b=True)
                if has_bias:
                    db = ops.colsum(dy, scale=alpha)
            dx = ops.gemm_tn(d16, w16, alpha=alpha, trans_b=True)
            wgrad_join(local=True)
        else:
            x, ws = ctx.saved_tensors[0], ctx.saved_tensors[1:]
            dx = ops.gemm_tn(operand(dy, 'a'), concat_kT_operand(list(ws), 'b', dy.device), alpha=alpha)
            dW = ops.gemm_tn(operand(dy, 'a', transpose=True), operand(x, 'b', transpose=True),
                             alpha=alpha)
            db = ops.colsum(dy, scale=alpha) if has_bias else None
        dws = tuple(dW[c * E:(c + 1) * E] for c in range(n))
        dbs = (None,) * n
        if has_bias:
            dbs = tuple(db[c * E:(c + 1) * E] for c in range(n))
        return (dx, None, None) + dws + dbs


class MultiCtxAttentionFn(Function):
    """All n cross-attentions of a layer on one fused query buffer: q_c = Q_all[:, cE:(c+1)E],
    out_c written into out_all[:, cE:(c+1)E] (row strides do the slicing, no copies)."""

    @staticmethod
    def forward(ctx, q_all, T, B, H, zero_row, p, seeds, need_weights, n, *args):
        kvs, bks, bvs, masks = args[:n], args[n:2 * n], args[2 * n:3 * n], args[3 * n:4 * n]
        # optional (GradSlab, layer) per context: where the backward writes dL/dkv
        ctx.slabs = args[4 * n] if len(args) > 4 * n else (None,) * n
        # optional per-context valid key counts (int32 [B]): trailing padding is skipped
        ctx.kv_lens = args[4 * n + 1] if len(args) > 4 * n + 1 else (None,) * n
        E = q_all.shape[1] // n
        D = E // H
        tc = config.precision == 'bf16' and D == 64
        out_all = torch.empty_like(q_all)
        lses, weights, Ss = [], [], []
        multi = tc and n <= 4 and config.attn_multi      # one launch for all contexts of the layer
        tw = multi and _tw() and E % 8 == 0
        out16_all = torch.empty(q_all.shape, dtype=torch.bfloat16, device=q_all.device) if tw else None
        items = []
        for c in range(n):
            kv = kvs[c]
            S = kv.shape[0] // B if kv is not None else 0
            Ss.append(S)
            q = q_all[:, c * E:(c + 1) * E]
            bk = bks[c].view(-1) if bks[c] is not None else None
            bv = bvs[c].view(-1) if bvs[c] is not None else None
            if multi:
                lse = torch.empty((B, H, T), dtype=torch.float32, device=q_all.device)
                items.append(dict(q=q, k=kv[:, :E] if S > 0 else None, v=kv[:, E:] if S > 0 else None,
                                  bias_k=bk, bias_v=bv, mask=masks[c], out=out_all[:, c * E:(c + 1) * E],
                                  lse=lse, S=S, seed=seeds[c], kv_len=ctx.kv_lens[c],
                                  out16=out16_all[:, c * E:(c + 1) * E] if tw else None))
            else:
                _, lse = ops.attn_fwd(q, kv[:, :E] if S > 0 else None, kv[:, E:] if S > 0 else None, bk, bv,
                                      masks[c] if S > 0 else None, T, B, S, H, D, zero_row, p, seeds[c],
                                      tc=tc, out=out_all[:, c * E:(c + 1) * E])
            lses.append(lse)
        if multi:
            ops.attn_fwd_tc_multi(items, T, B, H, D, zero_row, p)
            if tw:
                twin.put(out_all, out16_all)        # operand of the out-projections
        for c in range(n):
            kv, S = kvs[c], Ss[c]
            q = q_all[:, c * E:(c + 1) * E]
            bk = bks[c].view(-1) if bks[c] is not None else None
            lse = lses[c]
            if need_weights:
                weights.append(ops.attn_avg_weights(q, kv[:, :E] if S > 0 else None, bk,
                                                    masks[c] if S > 0 else None, lse, T, B, S, H, D,
                                                    zero_row))
        ctx.cfg = (T, B, H, D, zero_row, p, seeds, n, tuple(Ss), tc, multi)
        ctx.save_for_backward(q_all, out_all, *kvs, *bks, *bvs, *masks, *lses)
        ctx.mark_non_differentiable(*weights)
        return (out_all,) + tuple(weights)

    @staticmethod
    def backward(ctx, dout_all, *_dw):
        T, B, H, D, zero_row, p, seeds, n, Ss, tc, multi = ctx.cfg
        sv = ctx.saved_tensors
        q_all, out_all = sv[0], sv[1]
        kvs, bks, bvs = sv[2:2 + n], sv[2 + n:2 + 2 * n], sv[2 + 2 * n:2 + 3 * n]
        masks, lses = sv[2 + 3 * n:2 + 4 * n], sv[2 + 4 * n:2 + 5 * n]
        E = H * D
        dout_all = _c(dout_all)
        dq_all = torch.empty_like(q_all)
        tw = multi and _tw() and E % 8 == 0
        dq16_all = torch.empty(q_all.shape, dtype=torch.bfloat16, device=q_all.device) if tw else None
        dkvs, dbks, dbvs = [], [], []
        items = []
        # D = rowsum(dO * O) of every context: written by the dQ kernel, read by the dK|dV kernel
        dsum = torch.empty((n, B * H * T), dtype=torch.float32, device=q_all.device) if multi else None
        for c in range(n):
            S, kv = Ss[c], kvs[c]
            if S > 0 and ctx.slabs[c] is not None:
                dkv = ctx.slabs[c][0].block(ctx.slabs[c][1])
            else:
                dkv = torch.empty_like(kv) if S > 0 else None
            dbk = ops.zeros_f32(bks[c].shape, bks[c]) if bks[c] is not None else None
            dbv = ops.zeros_f32(bvs[c].shape, bvs[c]) if bvs[c] is not None else None
            sl = slice(c * E, (c + 1) * E)
            if multi:
                items.append(dict(q=q_all[:, sl], k=kv[:, :E] if S > 0 else None, v=kv[:, E:] if S > 0 else None,
                                  bias_k=bks[c].view(-1) if bks[c] is not None else None,
                                  bias_v=bvs[c].view(-1) if bvs[c] is not None else None,
                                  mask=masks[c], out=out_all[:, sl], lse=lses[c], S=S, seed=seeds[c],
                                  dout=dout_all[:, sl], dq=dq_all[:, sl],
                                  dk=dkv[:, :E] if S > 0 else None, dv=dkv[:, E:] if S > 0 else None,
                                  dbias_k=dbk, dbias_v=dbv, kv_len=ctx.kv_lens[c],
                                  dq16=dq16_all[:, sl] if tw else None, dsum=dsum[c]))
            else:
                ops.attn_bwd(dout_all[:, sl], q_all[:, sl], kv[:, :E] if S > 0 else None,
                             kv[:, E:] if S > 0 else None,
                             bks[c].view(-1) if bks[c] is not None else None,
                             bvs[c].view(-1) if bvs[c] is not None else None,
                             masks[c] if S > 0 else None, out_all[:, sl], lses[c], dq_all[:, sl],
                             dkv[:, :E] if S > 0 else None, dkv[:, E:] if S > 0 else None, dbk, dbv,
                             T, B, S, H, D, zero_row, p, seeds[c], tc=tc)
            dkvs.append(dkv)
            dbks.append(dbk)
            dbvs.append(dbv)
        if multi:
            ops.attn_bwd_tc_multi(items, T, B, H, D, zero_row, p)
            if tw:
                twin.put(dq_all, dq16_all)          # operand of the query projection's backward GEMMs
        return (dq_all,) + (None,) * 8 + tuple(dkvs) + tuple(dbks) + tuple(dbvs) + (None,) * n \
            + (None,) * max(0, len(ctx.needs_input_grad) - (9 + 4 * n))


class FusedOutProjFn(Function):
    """h_c = a_c @ Wo_c^T + bo_c for the n attention outputs living side by side in a_all [N, n*E]
    (multi_head.py:476 out_proj).  Throughput mode: ONE batched launch for the n same-shape GEMMs
    (tt_gemm_bf16_tn nbatch: batch c reads column block c of the shared bf16 operand and row block c
    of the stacked weights), and likewise ONE launch for the n dX and ONE for the n dW GEMMs of the
    backward -- 3 launches per layer instead of 12.  Parity mode: n GEMMs on re-cast operands."""

    @staticmethod
    def forward(ctx, a_all, n, *args):
        ws, bs = args[:n], args[n:2 * n]
        E = a_all.shape[1] // n
        N = a_all.shape[0]
        fast = _fast()
        batched = fast and config.gemm_batched and n > 1 and E % 64 == 0 and n <= 16 \
            and all(w.shape == (E, E) for w in ws)
        has_bias = bs[0] is not None
        if batched:
            a16 = operand(a_all, 'a')
            w16 = concat_rows_operand(list(ws), 'b', a_all.device)          # [n*E, E], one prep launch
            bias = torch.cat([b.reshape(-1) for b in bs]) if has_bias else None
            out = torch.empty((n, N, E), dtype=torch.float32, device=a_all.device)
            ops.gemm_tn_batched(a16, w16, n, N, E, E, out.view(n * N, E), N * E, a_off=(E, 0), b_off=(0, E),
                                bias=bias, bias_off=E)
            ctx.cfg = (n, E, fast, has_bias, True)
            ctx.save_for_backward(a16, w16)
            return tuple(out[c] for c in range(n))
        a16 = operand(a_all, 'a')
        rep = _rep()
        w16s, hs = [], []
        for c in range(n):
            w16 = operand(ws[c], 'b')
            w16s.append(w16)
            if rep == 1:
                a_c = a16[:, c * E:(c + 1) * E]
            else:   # split layout is per full row: cast the block on its own
                a_c = operand(a_all[:, c * E:(c + 1) * E], 'a')
            hs.append(ops.gemm_tn(a_c, w16, bias=bs[c]))
        ctx.cfg = (n, E, fast, has_bias, False)
        if fast:
            ctx.save_for_backward(a16, *w16s)
        else:
            ctx.save_for_backward(a_all, *ws)
        return tuple(hs)

    @staticmethod
    def backward(ctx, *dhs):
        n, E, fast, has_bias, batched = ctx.cfg
        sv = ctx.saved_tensors
        a, ws = sv[0], sv[1:]
        N = a.shape[0]
        da_all = torch.empty((N, n * E), dtype=torch.float32, device=a.device)
        if batched:
            w16 = ws[0]                                                   # stacked [n*E, E]
            d16 = torch.empty((N, n * E), dtype=torch.bfloat16, device=a.device)
            for c in range(n):
                ops.cast_bf16(_c(dhs[c]), out=d16[:, c * E:(c + 1) * E])
            # dX_c = dh_c . W_c      (B stored [K = E_out, N = E_in]: row block c of the stack)
            ops.gemm_tn_batched(d16, w16, n, N, E, E, da_all, E, a_off=(E, 0), b_off=(0, E), trans_b=True)
            # dW_c = dh_c^T . a_c    (both operands stored [K = rows, .]: column blocks c)
            dW = torch.empty((n * E, E), dtype=torch.float32, device=a.device)
            ops.gemm_tn_batched(d16, a, n, E, E, N, dW, E * E, a_off=(E, 0), b_off=(E, 0), trans_a=True,
                                trans_b=True)
            dws = tuple(dW[c * E:(c + 1) * E] for c in range(n))
            dbs = (None,) * n
            if has_bias:
                db = ops.colsum(d16)                                      # one pass over the bf16 operand
                dbs = tuple(db[c * E:(c + 1) * E] for c in range(n))
            return (da_all, None) + dws + dbs
        dws, dbs = [], []
        for c in range(n):
            dh = _c(dhs[c])
            sl = slice(c * E, (c + 1) * E)
            if fast:
                d16 = operand(dh, 'a')
                with wgrad(d16, a, dh, local=True):    # dW_c, db_c alongside the dx GEMMs
                    dws.append(ops.gemm_tn(d16, a[:, sl], trans_a=True, trans_b=True))
                    dbs.append(ops.colsum(dh) if has_bias else None)
                ops.gemm_tn(d16, ws[c], out=da_all[:, sl], trans_b=True)
            else:
                ops.gemm_tn(operand(dh, 'a'), operand(ws[c], 'b', transpose=True), out=da_all[:, sl])
                dws.append(ops.gemm_tn(operand(dh, 'a', transpose=True),
                                       operand(a[:, sl], 'b', transpose=True)))
                dbs.append(ops.colsum(dh) if has_bias else None)
        wgrad_join(local=True)
        return (da_all, None) + tuple(dws) + tuple(dbs)


class OutProjContextLNFn(Function):
    """FusedOutProjFn + ContextLayerNormFn as one autograd node (throughput mode): the n attention
    outputs side by side in a_all [N, n*E] go through their out-projections (ONE batched GEMM launch,
    multi_head.py:476), then Y[:, cE:(c+1)E] = LN_c(X + dropout_c(h_c)) for all contexts in ONE launch
    that also writes the bf16 operand of context_fc (decoder_faces_objects.py:272-354).  Backward: ONE
    LayerNorm launch that writes dL/dh of all contexts straight as the bf16 operand of the batched dX
    and dW GEMMs and sums the residual gradients; no fp32 dL/dh is materialised."""

    @staticmethod
    def usable(a_all, n, ws):
        E = a_all.shape[1] // n
        return _fast() and _tw() and config.gemm_batched and 1 < n <= 4 and E % 64 == 0 and E <= 1024 \
            and all(w.shape == (E, E) for w in ws)

    @staticmethod
    def forward(ctx, a_all, X, p, seeds, eps, n, *args):
        ws, bs, gammas, betas = args[:n], args[n:2 * n], args[2 * n:3 * n], args[3 * n:4 * n]
        N, E = X.shape
        a16 = operand(a_all, 'a')
        w16 = concat_rows_operand(list(ws), 'b', a_all.device)          # [n*E, E]
        has_bias = bs[0] is not None
        bias = torch.cat([b.reshape(-1) for b in bs]) if has_bias else None
        hs = torch.empty((n, N, E), dtype=torch.float32, device=a_all.device)
        ops.gemm_tn_batched(a16, w16, n, N, E, E, hs.view(n * N, E), N * E, a_off=(E, 0), b_off=(0, E),
                            bias=bias, bias_off=E)
        Y, Y16, means, rstds = ops.ln_fwd_multi([hs[c] for c in range(n)], _c(X), gammas, betas, eps, p,
                                                seeds)
        twin.put(Y, Y16)
        ctx.cfg = (n, E, p, seeds, has_bias)
        ctx.save_for_backward(a16, w16, hs, *means, *rstds, *gammas)
        return Y

    @staticmethod
    def backward(ctx, dY):
        n, E, p, seeds, has_bias = ctx.cfg
        sv = ctx.saved_tensors
        a16, w16, hs = sv[:3]
        means, rstds, gammas = sv[3:3 + n], sv[3 + n:3 + 2 * n], sv[3 + 2 * n:3 + 3 * n]
        N = hs.shape[1]
        dev = dY.device
        dgb = ops.zeros_f32((2 * n, E), dY)
        dX, _, d16 = ops.ln_bwd_multi(_c(dY), [hs[c] for c in range(n)], means, rstds, gammas, p, seeds,
                                      want_dx=True, want_dh32=False, want_dh16=True,
                                      dgammas=[dgb[c] for c in range(n)],
                                      dbetas=[dgb[n + c] for c in range(n)])
        da_all = torch.empty((N, n * E), dtype=torch.float32, device=dev)
        ops.gemm_tn_batched(d16, w16, n, N, E, E, da_all, E, a_off=(E, 0), b_off=(0, E), trans_b=True)
        dW = torch.empty((n * E, E), dtype=torch.float32, device=dev)
        ops.gemm_tn_batched(d16, a16, n, E, E, N, dW, E * E, a_off=(E, 0), b_off=(E, 0), trans_a=True,
                            trans_b=True)
        dws = tuple(dW[c * E:(c + 1) * E] for c in range(n))
        dbs = (None,) * n
        if has_bias:
            db = ops.colsum(d16)
            dbs = tuple(db[c * E:(c + 1) * E] for c in range(n))
        return (da_all, dX, None, None, None, None) + dws + dbs + tuple(dgb[c] for c in range(n)) + \
            tuple(dgb[n + c] for c in range(n))


def _rep():
    return 1 if config.precision == 'bf16' else 3


def concat_k_operand(mats, side, device):
    """One bf16 operand whose K axis is the concatenation of `mats` ([R, K_i] each)."""
    R = mats[0].shape[0]
    Ktot = sum(m.shape[1] for m in mats)
    rep = _rep()
    if rep == 1 and side == 'b' and weight_bank.ACTIVE is not None:
        hit = weight_bank.ACTIVE.group('kcat', mats)
        if hit is not None:
            return hit
    buf = ops.bf16_buffer(R, Ktot * rep, device)
    split = 0 if rep == 1 else (1 if side == 'a' else 2)
    off = 0
    for m in mats:
        ops.cast_bf16(m, split=split, out=buf[:, off:], seg_stride=Ktot)
        off += m.shape[1]
    return buf


def concat_kT_operand(mats, side, device):
    """Operand [C, sum_i R_i] = concatenation along K of the TRANSPOSES of `mats` ([R_i, C])."""
    C = mats[0].shape[1]
    Ktot = sum(m.shape[0] for m in mats)
    rep = _rep()
    buf = ops.bf16_buffer(C, Ktot * rep, device)
    split = 0 if rep == 1 else (1 if side == 'a' else 2)
    off = 0
    for m in mats:
        ops.cast_bf16(m, transpose=True, split=split, out=buf[:, off:], seg_stride=Ktot)
        off += m.shape[0]
    return buf


class EmbedFn(Function):
    """AdaptiveEmbedding + sinusoidal positions (adaptive.py:61-76, positional.py:167-211,
    sum_text_field_embedder.py:117): out[t*B+b] = scale * sum_band proj_band(emb_band[id]) + pos.
    The band projections are one GEMM over K = n_bands*E (bands are disjoint, so the gathered
    operand is zero outside the token's band)."""

    @staticmethod
    def forward(ctx, ids, pos_table, start_pos, pad, scale, cutoffs, n, *args):
        # n may be (n_bands, padding_idx of the adaptive embedder): nn.Embedding(padding_idx) keeps
        # that local row of every band out of the gradient (adaptive.py:35-40)
        n, ctx.emb_pad = n if isinstance(n, tuple) else (n, 0)
        tables, projs = args[:n], args[n:2 * n]
        B, T = ids.shape
        E = projs[0].shape[0]
        A = ops.embed_gather(ids, cutoffs, tables, E, tbc=True)          # [T*B, n*E]
        pos = ops.make_positions(ids, pad, False, start_pos, tbc=True)   # [T,B] int32
        posv = ops.gather_rows(pos_table, pos.view(-1))                  # [T*B, E]
        w16 = concat_k_operand(list(projs), 'b', ids.device)             # [E, n*E(*3)]
        a16 = operand(A, 'a')
        out = ops.gemm_tn(a16, w16, alpha=scale, residual=posv)
        ctx.cfg = (n, scale, cutoffs, E, _fast())
        if ctx.cfg[4]:
            ctx.save_for_backward(ids, a16, w16, *tables)
        else:
            ctx.save_for_backward(ids, A, *tables, *projs)
        return out

    @staticmethod
    def backward(ctx, dout):
        n, scale, cutoffs, E, fast = ctx.cfg
        saved = ctx.saved_tensors
        dout = _c(dout)
        ids = saved[0]
        if fast:
            a16, w16 = saved[1], saved[2]
            tables = saved[3:3 + n]
            d16 = operand(dout, 'a')
            dA = ops.gemm_tn(d16, w16, alpha=scale, trans_b=True)                     # [N, n*E]
            dW = ops.gemm_tn(d16, a16, alpha=scale, trans_a=True, trans_b=True)       # [E, n*E]
            dprojs = [dW[:, i * E:(i + 1) * E] for i in range(n)]
        else:
            A = saved[1]
            tables, projs = saved[2:2 + n], saved[2 + n:2 + 2 * n]
            d16 = operand(dout, 'a')
            d16T = operand(dout, 'a', transpose=True)
            dA = torch.empty_like(A)
            dprojs = []
            for i in range(n):
                ops.gemm_tn(d16, operand(projs[i], 'b', transpose=True),
                            out=dA[:, i * E:(i + 1) * E], alpha=scale)
                dprojs.append(ops.gemm_tn(d16T, operand(A[:, i * E:(i + 1) * E], 'b', transpose=True),
                                          alpha=scale))
        dtables, ret = [], []
        for t in tables:
            g = _TIED_GRADS.pop(t.data_ptr(), None)
            if g is not None and g.shape == t.shape and g.is_contiguous() and g.dtype == torch.float32:
                dtables.append(g)        # tied: add into the softmax's gradient of the same matrix
                ret.append(None)
            else:
                g = torch.zeros_like(t)
                dtables.append(g)
                ret.append(g)
        _TIED_GRADS.clear()
        ops.embed_scatter_grad(ids, cutoffs, dtables, E, dA, padding_idx=ctx.emb_pad, tbc=True)
        return (None,) * 7 + tuple(ret) + tuple(dprojs)


class LayerMixFn(Function):
    """X_article = sum_l softmax(bert_weight)[l] * h_l  (transformer_faces_objects.py:355-364).
    hiddens: bf16 [L, R, E] from the frozen encoder (no grad); grad flows to bert_weight only."""

    @staticmethod
    def forward(ctx, hiddens, w):
        ctx.save_for_backward(hiddens, w)
        return ops.layer_mix_fwd(hiddens, w)

    @staticmethod
    def backward(ctx, dout):
        hiddens, w = ctx.saved_tensors
        return None, ops.layer_mix_bwd(hiddens, w, _c(dout))


class AdaptiveLossFn(Function):
    """AdaptiveSoftmax.forward + AdaptiveLoss + loss/ln2/ntokens in one node
    (softmax.py:169-191, adaptive_loss.py:27-73, transformer_faces_objects.py:82-90).
    Inputs: X [N,E]; target int64 [N]; word0 [c0,E], class_proj [n_tails,E],
    then per tail: proj_i [E,E], words_i [V_i,E].  Returns (loss [1], ntokens int32 [1]).
    Tail rows are compacted on the device (no nonzero() sync); every cluster's logits live in an
    fp32 buffer of capacity N rows and GEMMs stop at the device-side row count."""

    @staticmethod
    def forward(ctx, X, target, cutoffs, pad_idx, word0, class_proj, *tails):
        N, E = X.shape
        nt = len(cutoffs) - 1
        fast = _fast()
        head_t, tail_idx, tail_local, tail_count, ntok = ops.adaptive_prepare(target, cutoffs,
                                                                              pad_idx)
        hw16 = concat_rows_operand([word0, class_proj], 'b', X.device)
        x16 = operand(X, 'a')
        head_logits = ops.gemm_tn(x16, hw16, out=ops.f32_padded(N, hw16.shape[0], X))
        row_loss = torch.empty((nt + 1, N), dtype=torch.float32, device=X.device)
        head_lse, _ = ops.ce_fwd(head_logits, head_t, None, pad_idx, row_loss[0])
        saved_tail = []
        for i in range(nt):
            proj, words = tails[2 * i], tails[2 * i + 1]
            cnt = tail_count[i:i + 1]
            Xg = ops.gather_rows(X, tail_idx[i], cnt, cap=N)
            xg16, proj16 = operand(Xg, 'a'), operand(proj, 'b')
            P = ops.gemm_tn(xg16, proj16, m_limit=cnt)
            p16, words16 = operand(P, 'a'), operand(words, 'b')
            # rows >= cnt of the logits are never read (ce_fwd stops at cnt, ce_bwd zeroes them):
            # no zero fill of the [N, V_tail] buffer
            logits = ops.gemm_tn(p16, words16, m_limit=cnt, out=ops.f32_padded(N, words.shape[0], X))
            lse, _ = ops.ce_fwd(logits, tail_local[i], cnt, pad_idx, row_loss[i + 1])
            saved_tail += ([xg16, p16, proj16, words16, logits, lse] if fast
                           else [Xg, P, proj, words, logits, lse])
        loss, scale = ops.loss_finalize(row_loss, ntok)
        ctx.cfg = (cutoffs, pad_idx, nt, fast)
        ctx.tied_ptrs = (word0.data_ptr(),) + tuple(tails[2 * i + 1].data_ptr() for i in range(nt))
        head_saved = (x16, hw16) if fast else (X, word0, class_proj)
        ctx.n_head = len(head_saved)
        ctx.save_for_backward(*head_saved, head_logits, head_lse, head_t, tail_idx, tail_local,
                              tail_count, scale, *saved_tail)
        ctx.mark_non_differentiable(ntok)
        return loss, ntok

    @staticmethod
    def backward(ctx, dloss, _dntok):
        cutoffs, pad_idx, nt, fast = ctx.cfg
        s = ctx.saved_tensors
        head_saved, s = s[:ctx.n_head], s[ctx.n_head:]
        head_logits, head_lse, head_t, tail_idx, tail_local, tail_count, scale = s[:7]
        saved_tail = s[7:]
        c0 = cutoffs[0]
        gscale = ops.scalar_mul(scale, _c(dloss).view(1))       # upstream grad stays on device
        if fast:        # gradient written directly as the bf16 operand (no fp32 pass + cast)
            dlog16 = ops.ce_bwd16(head_logits, head_t, head_lse, gscale, None, pad_idx)
        else:
            dlog = ops.ce_bwd_(head_logits, head_t, head_lse, gscale, None, pad_idx)   # in place
            dlog16 = operand(dlog, 'a')
        if fast:
            x16, hw16 = head_saved
            dX = ops.gemm_tn(dlog16, hw16, trans_b=True)
            dW_head = ops.gemm_tn(dlog16, x16, trans_a=True, trans_b=True)
        else:
            X, word0, class_proj = head_saved
            dX = ops.gemm_tn(dlog16, concat_kT_operand([word0, class_proj], 'b', dlog.device))
            dW_head = ops.gemm_tn(operand(dlog, 'a', transpose=True), operand(X, 'b', transpose=True))
        dtails = []
        for i in range(nt):
            a0, a1, a2, a3, logits, lse = saved_tail[6 * i:6 * i + 6]
            cnt = tail_count[i:i + 1]
            if fast:    # rows >= round_up(cnt, 128) are never read by the cnt-limited GEMMs below
                dl16 = ops.ce_bwd16(logits, tail_local[i], lse, gscale, cnt, pad_idx, zero_round=128)
            else:
                dl = ops.ce_bwd_(logits, tail_local[i], lse, gscale, cnt, pad_idx)   # rows >= cnt := 0
                dl16 = operand(dl, 'a')
            if fast:
                xg16, p16, proj16, words16 = a0, a1, a2, a3
                # few rows x the whole tail vocabulary as the contraction: split-K (m_hint = the
                # expected row count lets the dispatcher see how few tiles there really are)
                dP = ops.gemm_tn(dl16, words16, m_limit=cnt, trans_b=True, m_hint=max(128, dl16.shape[0] // 8))
                # contraction over the cluster's rows: rows >= cnt of dl / dP are zero, so the K loop
                # stops at the device-side row count (23-120 of the 800 rows for the tail clusters)
                dwords = ops.gemm_tn(dl16, p16, trans_a=True, trans_b=True, k_limit=cnt)
                dP16 = operand(dP, 'a')
                dXg = ops.gemm_tn(dP16, proj16, m_limit=cnt, trans_b=True)
                dproj = ops.gemm_tn(dP16, xg16, trans_a=True, trans_b=True, k_limit=cnt)
            else:
                Xg, P, proj, words = a0, a1, a2, a3
                dP = ops.gemm_tn(dl16, operand(words, 'b', transpose=True), m_limit=cnt)
                dwords = ops.gemm_tn(operand(dl, 'a', transpose=True), operand(P, 'b', transpose=True))
                dXg = ops.gemm_tn(operand(dP, 'a'), operand(proj, 'b', transpose=True), m_limit=cnt)
                dproj = ops.gemm_tn(operand(dP, 'a', transpose=True), operand(Xg, 'b', transpose=True))
            ops.scatter_add_rows(dXg, tail_idx[i], dX, cnt)
            dtails += [dproj, dwords]
        _TIED_GRADS.clear()
        _TIED_GRADS[ctx.tied_ptrs[0]] = dW_head[:c0]
        for i in range(nt):
            _TIED_GRADS[ctx.tied_ptrs[1 + i]] = dtails[2 * i + 1]
        return (dX, None, None, None, dW_head[:c0], dW_head[c0:]) + tuple(dtails)
