"""Top-level caption models of tell/models/transformer_faces_objects.py and
tell/models/transformer_flattened.py on the B200 kernels.

Same registry names, constructor kwargs, forward()/generate() signatures and output dict keys as
the reference.  The two frozen encoders cannot be downloaded here (no network), so they are
constructed randomly initialised (or injected through the extra `resnet=` / `roberta=` kwargs);
their state dicts use torchvision's / fairseq's key names so real checkpoints load unchanged."""
import math

import os

import torch
import torch.nn as nn

from .. import config
from .. import functional as Fn
from .. import ops
from ..modules import Criterion
from ..registry import Registrable
from .decoder import Decoder
from .resnet import resnet152
from .roberta import RobertaEncoder


_SIDE = {}


def _side_stream(device):
    key = (device.type, device.index)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=device)
    return _SIDE[key]


class Model(nn.Module, Registrable):
    pass


class _CaptionModelBase(Model):
    USES_FACES = True
    USES_OBJECTS = True
    USES_IMAGE = True

    def __init__(self, vocab, decoder, criterion, evaluate_mode=False, attention_dim=1024,
                 hidden_size=1024, dropout=0.1, vocab_size=50264, model_name='roberta-base',
                 namespace='bpe', index='roberta', padding_value=1, use_context=True,
                 sampling_topk=1, sampling_temp=1.0, weigh_bert=False, initializer=None,
                 resnet=None, roberta=None):
        super().__init__()
        self.decoder = decoder
        self.criterion = criterion
        self.index = index
        self.namespace = namespace
        self.resnet = resnet if resnet is not None else resnet152()
        self.roberta = roberta if roberta is not None else RobertaEncoder()
        self.use_context = use_context
        self.padding_idx = padding_value
        self.evaluate_mode = evaluate_mode
        self.sampling_topk = sampling_topk
        self.sampling_temp = sampling_temp
        self.weigh_bert = weigh_bert
        n_hidden = self.roberta.n_layers + 1
        if weigh_bert:
            self.bert_weight = nn.Parameter(torch.empty(n_hidden).uniform_())
        for p in self.resnet.parameters():           # `no_grad: ^resnet|^roberta` (config.yaml:150)
            p.requires_grad_(False)
        for p in self.roberta.parameters():
            p.requires_grad_(False)
        self.n_batches = 0
        self.n_samples = 0
        self.gen_len = 100

    @classmethod
    def _from_params(cls, params, **extras):
        params = dict(params)
        decoder = Decoder.from_params(params.pop('decoder'))
        criterion = Criterion.from_params(params.pop('criterion'))
        params.pop('initializer', None)
        return cls(vocab=None, decoder=decoder, criterion=criterion, **params, **extras)

    # ------------------------------------------------------------------ frozen encoders
    @torch.no_grad()
    def encode(self, context, image, n_real_tokens=0):
        """The gradient-free part of _forward (:332, :352-353): ResNet features (NHWC bf16) and all
        RoBERTa hidden states (bf16 [L+1, B*S, E]).  It depends on no trainable weight, so a
        data-parallel trainer may run it for step i+1 while step i's gradient all-reduce is still
        in flight; pass the result to forward(..., encoded=...)."""
        with config.gemm_sm_cap(config.encoder_sm_cap):
            return self._encode(context, image, n_real_tokens)

    def _encode(self, context, image, n_real_tokens=0):
        if self.USES_IMAGE and config.encoder_overlap and image.is_cuda:
            # the two frozen encoders are independent: ResNet's ~200 small latency-bound launches
            # run as a parallel branch (one fork, one join) beside RoBERTa's large GEMMs
            cur = torch.cuda.current_stream(image.device)
            side = _side_stream(image.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                feats = self.resnet.features_nhwc(image)
            hid, _ = self.roberta.all_hiddens(context[self.index], n_real_tokens)
            cur.wait_stream(side)
            return feats, hid
        with config.nvtx_range('tt/encode'):
            feats = self.resnet.features_nhwc(image) if self.USES_IMAGE else None
            hid, _ = self.roberta.all_hiddens(context[self.index], n_real_tokens)
        return feats, hid

    # ------------------------------------------------------------------ _forward (:311-397)
    def _forward(self, context, image, caption, face_embeds=None, obj_embeds=None, encoded=None):
        caption_ids = caption[self.index]
        target_ids = caption_ids[:, 1:].contiguous()
        caption_ids = caption_ids[:, :-1].contiguous()
        caption[self.index] = caption_ids            # the reference mutates the dict (:329)
        article_ids = context[self.index]
        B, S = article_ids.shape
        contexts = {}
        if encoded is None:
            encoded = self.encode(context, image)
        feats, hid = encoded
        if self.USES_IMAGE:
            P = feats.shape[1] * feats.shape[2]                      # [B,7,7,2048] bf16 == [B,49,2048]
            X_image = ops.bf16_to_f32(feats).view(B, P, feats.shape[3])
            contexts['image'] = Fn.Transpose01Fn.apply(X_image)      # [49,B,2048]
            contexts['image_mask'] = torch.zeros((B, P), dtype=torch.bool, device=feats.device)
        if self.weigh_bert:                                          # hid: bf16 [25, B*S, E]
            X_article = Fn.LayerMixFn.apply(hid, self.bert_weight)
        else:
            X_article = ops.bf16_to_f32(hid[-1])
        E = X_article.shape[-1]
        contexts['article'] = Fn.Transpose01Fn.apply(X_article.view(B, S, E))
        contexts['article_mask'] = article_ids == self.padding_idx
        contexts['sections'] = None
        contexts['sections_mask'] = None
        if self.USES_FACES:
            face_masks = ops.nan_rows_(face_embeds)                  # in place, like the reference
            contexts['faces'] = Fn.Transpose01Fn.apply(face_embeds) if face_embeds.shape[2] > 0 \
                else face_embeds.transpose(0, 1)
            contexts['faces_mask'] = face_masks
        if self.USES_OBJECTS:
            obj_masks = ops.nan_rows_(obj_embeds)
            contexts['obj'] = Fn.Transpose01Fn.apply(obj_embeds) if obj_embeds.shape[2] > 0 \
                else obj_embeds.transpose(0, 1)
            contexts['obj_mask'] = obj_masks
        return caption_ids, target_ids, contexts

    # ------------------------------------------------------------------ forward (:67-140)
    def forward(self, context, image, caption, face_embeds=None, obj_embeds=None, metadata=None,
                names=None, attn_idx=None, encoded=None):
        if config.zero_arena is not None and torch.is_grad_enabled():
            config.zero_arena.reset()          # one memset for every small gradient of this step
        with config.nvtx_range('tt/_forward'):
            caption_ids, target_ids, contexts = self._forward(context, image, caption, face_embeds,
                                                              obj_embeds, encoded)
        with self.decoder.weight_scope():      # one launch prepares every decoder weight operand
            with config.nvtx_range('tt/decoder'):
                X, _ = self.decoder.forward_tbc(caption, contexts)       # [T,B,E], no transpose needed
            T, B, E = X.shape
            # the loss is a sum over tokens, so (t,b) order with the target transposed alike is exact
            tgt_tb = target_ids.t().contiguous()
            with config.nvtx_range('tt/adaptive_loss'):
                loss, ntokens = self.criterion.fused(self.decoder.adaptive_softmax,
                                                     (X.view(T * B, E), None), tgt_tb)
        output_dict = {'loss': loss.view(()), 'sample_size': ntokens}
        if not self.training and self.evaluate_mode:
            with config.nvtx_range('tt/_generate'):
                _, gen_ids, attns = self._generate(caption_ids, contexts, attn_idx)
            output_dict['captions'] = [m['caption'] for m in metadata] if metadata else None
            output_dict['metadata'] = metadata
            output_dict['attns'] = attns
            output_dict['gen_ids'] = gen_ids.cpu().numpy()
        self.n_samples += caption_ids.shape[0]
        self.n_batches += 1
        return output_dict

    # ------------------------------------------------------------------ _generate (:399-494)
    @torch.no_grad()
    @torch.no_grad()
    def _generate(self, caption_ids, contexts, attn_idx=None, early_exit=True, sync_every=8):
        """Greedy decode (sampling_topk=1: multinomial over one candidate == argmax).  Device
        resident: all rows stay in the batch, finished rows are masked (emit pad / log-prob 0),
        projected contexts are cached, and the host is consulted only every `sync_every` steps.
        Emits the same token_ids [B, 1+n] / log_probs [B, n] matrices as the reference's
        compacting loop."""
        topk = int(self.sampling_topk)
        if topk < 1:
            raise ValueError('sampling_topk must be >= 1')
        was_training = self.decoder.training
        self.decoder.eval()
        need_attn = [l.need_attn for l in self.decoder.layers]
        for l in self.decoder.layers:
            l.need_attn = False
        eos, pad = 2, self.padding_idx
        B = caption_ids.shape[0]
        dev = caption_ids.device
        if self.decode_graph and dev.type == 'cuda' and self.gen_len > 2 and topk == 1 \
                and not torch.cuda.is_current_stream_capturing():
            with torch.no_grad():
                log_probs, token_ids = self._generate_graphed(caption_ids, contexts, early_exit,
                                                              sync_every)
            for l, n in zip(self.decoder.layers, need_attn):
                l.need_attn = n
            self.decoder.train(was_training)
            return log_probs, token_ids, []
        state = {}
        prev = caption_ids[:, 0:1].contiguous()
        active = prev[:, 0] != eos
        ids_cols, lp_cols, any_active = [prev], [], []
        n_steps = 0
        for i in range(self.gen_len):
            # weights are constant during decoding: operands are prepared at step 0 only
            with self.decoder.weight_scope(refresh=(i == 0)):
                X, _ = self.decoder.forward_tbc({self.index: prev}, contexts,
                                                incremental_state=state)
                if topk == 1:
                    tok, lp = self.decoder.adaptive_softmax.greedy(X.view(B, -1))
                else:
                    # :450-464 -- top-k of the full-vocabulary log-probs, temperature, one multinomial
                    # draw among the k candidates (torch's RNG stream, as in the reference)
                    lprobs = self.decoder.adaptive_softmax.get_log_prob(X.view(B, 1, -1))[:, 0]
                    top_lp, top_idx = lprobs.topk(topk)
                    pick = torch.multinomial((top_lp / self.sampling_temp).exp(), num_samples=1)
                    tok = top_idx.gather(-1, pick).view(B)
                    lp = top_lp.gather(-1, pick).view(B)
            lp = lp / self.sampling_temp
            tok = torch.where(active, tok, torch.full_like(tok, pad))
            lp = torch.where(active, lp, torch.zeros_like(lp))
            ids_cols.append(tok.view(B, 1))
            lp_cols.append(lp.view(B, 1))
            active = active & (tok != eos)
            any_active.append(active.any())
            prev = tok.view(B, 1)
            n_steps += 1
            if early_exit and (i + 1) % sync_every == 0 and not bool(any_active[-1]):
                break
        if early_exit:
            flags = torch.stack(any_active).cpu()
            dead = (~flags).nonzero()
            if dead.numel() > 0:
                n_steps = int(dead[0]) + 1       # the reference stops after the first all-done step
        token_ids = torch.cat(ids_cols[:1 + n_steps], dim=1)
        log_probs = torch.cat(lp_cols[:n_steps], dim=1)
        for l, n in zip(self.decoder.layers, need_attn):
            l.need_attn = n
        self.decoder.train(was_training)
        return log_probs, token_ids, []

    decode_graph = True

    decode_graph_reuse = True      # keep the captured decode step across generate() calls of one shape
    # capture the decode step (one chain, nothing beside it) with programmatic dependent launch
    decode_pdl = os.environ.get('TT_DECODE_PDL', '1') == '1'

    def _decode_setup(self, caption_ids, contexts):
        """Static device buffers of one greedy decode + the step function that updates them in place."""
        from ..modules.token_embedders import POSITION_DEV_KEY
        eos, pad = 2, self.padding_idx
        B = caption_ids.shape[0]
        dev = caption_ids.device
        n_max = self.gen_len
        dec = self.decoder

        class Run:
            pass
        r = Run()
        r.contexts = contexts
        r.ids_buf = torch.full((B, 1 + n_max), pad, dtype=torch.long, device=dev)
        r.lp_buf = torch.zeros((B, n_max), dtype=torch.float32, device=dev)
        r.flags = torch.zeros(n_max, dtype=torch.bool, device=dev)
        r.prev = caption_ids[:, 0:1].contiguous().clone()
        r.ids_buf[:, 0:1] = r.prev
        r.active = r.prev[:, 0] != eos
        r.pos = torch.zeros(1, dtype=torch.int32, device=dev)          # tokens fed so far
        r.col = torch.zeros(1, dtype=torch.long, device=dev)           # output column of this step
        r.state = {POSITION_DEV_KEY: r.pos}

        def one_step():
            X, _ = dec.forward_tbc({self.index: r.prev}, r.contexts, incremental_state=r.state)
            tok, lp = dec.adaptive_softmax.greedy(X.view(B, -1))
            lp = lp / self.sampling_temp
            tok = torch.where(r.active, tok, torch.full_like(tok, pad))
            lp = torch.where(r.active, lp, torch.zeros_like(lp))
            r.ids_buf.index_copy_(1, r.col + 1, tok.view(B, 1))
            r.lp_buf.index_copy_(1, r.col, lp.view(B, 1))
            r.active.logical_and_(tok != eos)
            r.flags.index_copy_(0, r.col, r.active.any().view(1))
            r.prev.copy_(tok.view(B, 1))
            r.pos.add_(1)
            r.col.add_(1)
        r.one_step = one_step
        return r

    @staticmethod
    def _copy_tree(dst, src):
        """dst <- src for two structurally identical trees of tensors (dicts / lists / tuples); False when
        the structure, a shape or a dtype differs (the caller then captures a new graph)."""
        if isinstance(dst, torch.Tensor) or isinstance(src, torch.Tensor):
            if not (isinstance(dst, torch.Tensor) and isinstance(src, torch.Tensor)) or \
                    dst.shape != src.shape or dst.dtype != src.dtype or dst.stride() != src.stride():
                return False
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src)
            return True
        if isinstance(dst, dict):
            if not isinstance(src, dict) or dst.keys() != src.keys():
                return False
            return all(_CaptionModelBase._copy_tree(dst[k], src[k]) for k in dst)
        if isinstance(dst, (list, tuple)):
            if not isinstance(src, (list, tuple)) or len(dst) != len(src):
                return False
            return all(_CaptionModelBase._copy_tree(d, s_) for d, s_ in zip(dst, src))
        if isinstance(dst, (set, frozenset)):
            return isinstance(src, (set, frozenset)) and len(dst) == len(src)
        if dst is None or src is None:
            return dst is None and src is None
        try:
            return bool(dst == src)              # scalar state (batch size, flags) must agree
        except Exception:
            return False

    def _generate_graphed(self, caption_ids, contexts, early_exit, sync_every):
        """The greedy loop with ONE captured decode step replayed gen_len-1 times.

        Step 0 runs eagerly (it projects the four contexts, allocates the fixed-size DynamicConv
        input buffers and prepares the weight operands); every later step has identical shapes, so
        a single CUDA graph holds it: previous token, finished-row mask, running position and the
        output matrices are static device buffers the graph updates in place (the column written
        is indexed by a device step counter).  The host only replays, and looks at the all-done
        flag every `sync_every` steps.

        The captured step is KEPT across calls (decode_graph_reuse): a later call with the same shapes
        runs its step 0 on fresh buffers, copies the resulting state (head-major K|V caches,
        DynamicConv windows, masks, counters) into the buffers the graph was captured on and replays --
        no second eager step, no capture, no cudaGraphInstantiate (10-140 ms of host time per call,
        against ~65 ms of device time for 48 replays at batch 256)."""
        B = caption_ids.shape[0]
        dev = caption_ids.device
        n_max = self.gen_len
        dec = self.decoder
        for m in dec.modules():                                      # positional table for all steps
            if hasattr(m, 'ensure_size') and hasattr(m, 'padding_idx'):
                m.ensure_size(n_max + 2 + m.padding_idx)
        timing = getattr(self, 'decode_timing', None)     # optional dict: CUDA events around the phases
        if timing is not None:
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            evs[0].record()
        run = self._decode_setup(caption_ids, contexts)
        with dec.weight_scope(refresh=True):
            run.one_step()                                           # step 0, eager
        dec.build_decode_cache(run.state, contexts)   # head-major K|V (+ valid key counts) for the T = 1 attention kernel
        if timing is not None:
            evs[1].record()
        key = (B, n_max, float(self.sampling_temp), str(dev),
               tuple(sorted((k, tuple(v.shape), str(v.dtype)) for k, v in contexts.items()
                            if isinstance(v, torch.Tensor))))
        kept = getattr(self, '_decode_graph_keep', None)
        graph = None
        first_replay = 2
        if self.decode_graph_reuse and kept is not None and kept[0] == key:
            _, g_old, st = kept
            ok = self._copy_tree(st.state, run.state)
            ok = ok and self._copy_tree({k: v for k, v in st.contexts.items() if isinstance(v, torch.Tensor)},
                                        {k: v for k, v in contexts.items() if isinstance(v, torch.Tensor)})
            ok = ok and self._copy_tree([st.ids_buf, st.lp_buf, st.flags, st.prev, st.active, st.pos, st.col],
                                        [run.ids_buf, run.lp_buf, run.flags, run.prev, run.active, run.pos,
                                         run.col])
            if ok:
                graph, run, first_replay = g_old, st, 1      # the graph's own buffers now hold this call
        if graph is None:
            # Step 1 runs eagerly on the capture stream (warms every lazy path), then the same stream
            # records the step.  capture_begin/capture_end directly: torch.cuda.graph() would also
            # synchronise the device, run the Python GC and empty the allocator cache on every call.
            cur = torch.cuda.current_stream(dev)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(cur)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side), dec.weight_scope(refresh=False), config.pdl(self.decode_pdl):
                run.one_step()
                # one allocator pool for every decode graph of this model: the blocks of a replaced
                # graph are reused by the next capture instead of cudaMalloc / cudaFree per capture
                if kept is not None:
                    graph.capture_begin(pool=kept[1].pool())
                else:
                    graph.capture_begin()
                try:
                    run.one_step()                   # recorded, not executed: later steps are replays
                finally:
                    graph.capture_end()
            cur.wait_stream(side)
            object.__setattr__(self, '_decode_graph_keep', (key, graph, run))
        if timing is not None:
            evs[2].record()
        n_steps = n_max
        flags = run.flags
        for i in range(first_replay, n_max):
            graph.replay()
            if early_exit and (i + 1) % sync_every == 0 and not bool(flags[i]):
                n_steps = i + 1
                break
        if timing is not None:
            evs[3].record()
            torch.cuda.synchronize()
            timing.update(step0_and_cache_ms=evs[0].elapsed_time(evs[1]),
                          step1_and_capture_ms=evs[1].elapsed_time(evs[2]),
                          replays_ms=evs[2].elapsed_time(evs[3]), replays=n_steps - first_replay,
                          graph_reused=first_replay == 1)
        if early_exit:
            f = flags[:n_steps].cpu()
            dead = (~f).nonzero()
            if dead.numel() > 0:
                n_steps = int(dead[0]) + 1       # the reference stops after the first all-done step
        return run.lp_buf[:, :n_steps].clone(), run.ids_buf[:, :1 + n_steps].clone()

    # ------------------------------------------------------------------ generate (:142-309)
    @torch.no_grad()
    def generate(self, context, image, face_embeds=None, obj_embeds=None, metadata=None,
                 names=None, attn_idx=None):
        """Demo entry point: same inputs as the reference; returns generated ids and log-probs
        (the word-level attention post-processing of :159-304 is host-side string work, out of
        scope)."""
        B = image.shape[0]
        caption = {self.index: context[self.index].new_zeros(B, 2)}
        caption_ids, _, contexts = self._forward(context, image, caption, face_embeds, obj_embeds)
        log_probs, gen_ids, _ = self._generate(caption_ids, contexts, attn_idx)
        return {'generated_indices': gen_ids, 'log_probs': log_probs, 'metadata': metadata}

    def get_metrics(self, reset=False):
        metrics = {'_n_batches': self.n_batches, '_n_samples': self.n_samples}
        if reset:
            self.n_batches = 0
            self.n_samples = 0
        return metrics


@Model.register('transformer_faces_objects')
class TransformerFacesObjectModel(_CaptionModelBase):
    pass


@Model.register('transformer_faces')
class TransformerFacesModel(_CaptionModelBase):
    USES_OBJECTS = False


@Model.register('transformer_flattened')
class TransformerFlattenedModel(_CaptionModelBase):
    """transformer_flattened.py:23-238 (used by 4_no_image .. 7_*): no faces / objects.  As in the
    reference, the ResNet still runs even when the decoder ignores the image
    (transformer_flattened.py:49,185)."""
    USES_FACES = False
    USES_OBJECTS = False
