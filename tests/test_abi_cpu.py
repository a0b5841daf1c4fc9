"""CPU suite: the C-ABI library loads and exports every symbol include/tt_b200.h declares; the
host-side registry builds the reference's YAML model blocks; no compute is launched."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    import sys
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build()


def test_library_exports_every_declared_symbol():
    _ensure_built()
    from tell_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'tt_b200.h')).read()
    names = set(re.findall(r'\b(tt_[a-z0-9_]+)\s*\(', hdr))
    assert len(names) >= 30
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    lib.tt_abi_version.restype = ctypes.c_int
    assert lib.tt_abi_version() == 1
    lib.tt_last_error.restype = ctypes.c_char_p
    assert lib.tt_last_error() is not None


def test_ctypes_argument_blocks_match_the_header(tmp_path):
    """Every struct that crosses the C ABI by pointer (TtGemmParams, TtAttnCtx, TtLnFwdMulti, TtLnBwdMulti,
    ...) has the size and the field offsets the header gives it: compiled from include/tt_b200.h with gcc
    and compared with the ctypes mirrors in tell_b200/_lib.py, so a field added on one side only fails
    here instead of corrupting arguments on the GPU."""
    import subprocess
    from tell_b200 import _lib, optim, weight_bank
    structs = {'TtGemmParams': _lib.TtGemmParams, 'TtAttnCtx': _lib.TtAttnCtx,
               'TtLnFwdMulti': _lib.TtLnFwdMulti, 'TtLnBwdMulti': _lib.TtLnBwdMulti,
               'TtPrepSeg': weight_bank.TtPrepSeg, 'TtWnormBwdSeg': weight_bank.TtWnormBwdSeg,
               'TtAdamSeg': optim.TtAdamSeg, 'TtAdamHyper': optim.TtAdamHyper}
    hdr = open(os.path.join(ROOT, 'include', 'tt_b200.h')).read()
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "tt_b200.h"', 'int main(void) {']
    for name, cls in structs.items():
        assert re.search(r'\}\s*%s\s*;' % name, hdr), '%s is not declared in the header' % name
        lines.append('  printf("%s size %%zu\\n", sizeof(%s));' % (name, name))
        for fname, _ in cls._fields_:
            lines.append('  printf("%s %s %%zu\\n", offsetof(%s, %s));' % (name, fname, name, fname))
    lines += ['  return 0;', '}']
    src = tmp_path / 'abi.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'abi'
    subprocess.run(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    got = {}
    for ln in out.splitlines():
        a, b, c = ln.split()
        got[(a, b)] = int(c)
    for name, cls in structs.items():
        assert got[(name, 'size')] == ctypes.sizeof(cls), (name, got[(name, 'size')], ctypes.sizeof(cls))
        for fname, _ in cls._fields_:
            assert got[(name, fname)] == getattr(cls, fname).offset, (name, fname)


def test_invalid_arguments_fail_loudly_without_a_gpu():
    _ensure_built()
    from tell_b200 import _lib
    lib = _lib.lib()
    rc = lib.tt_gemm_bf16_tn(None, None)
    assert rc == -1 and b'null params' in lib.tt_last_error()
    rc = lib.tt_glu_fwd(None, None, ctypes.c_longlong(4), ctypes.c_int(8), None)
    assert rc == -1
    # argument validation happens before any launch, so it is observable without a device
    ll, ci = ctypes.c_longlong, ctypes.c_int
    rc = lib.tt_ce_bwd_bf16(None, ll(8), None, None, ci(4), ci(8), ci(1), None, None, None, ll(8),
                            ci(0), None)
    assert rc == -1 and b'tt_ce_bwd_bf16' in lib.tt_last_error()
    rc = lib.tt_im2col_nhwc(None, None, ci(1), ci(4), ci(4), ci(8), ci(3), ci(3), ci(1), ci(1), ci(72),
                            None)
    assert rc == -1 and b'tt_im2col_nhwc' in lib.tt_last_error()
    rc = lib.tt_cast_bf16(None, ll(8), None, ll(8), ci(4), ci(8), ci(0), ci(0), ll(0), None)
    assert rc == -1 and b'tt_cast_bf16' in lib.tt_last_error()


def test_ops_refuse_cpu_tensors():
    import torch
    from tell_b200 import _lib, ops
    with pytest.raises(_lib.TtError):
        ops.cast_bf16(torch.zeros(4, 8))


def test_registry_builds_reference_model_block():
    """The `model:` block of expt/nytimes/9_transformer_objects/config.yaml (decoder + criterion
    part, keys verbatim) instantiates through the registry; unknown keys are rejected."""
    from tell_b200.models.decoder import Decoder
    from tell_b200.registry import ConfigurationError
    import tell_b200.models  # noqa: F401
    block = {
        'type': 'dynamic_conv_decoder_faces_objects',
        'embedder': {'type': 'sum', 'token_embedders': {
            'adaptive': {'type': 'adaptive', 'namespace': 'bpe', 'padding_idx': 0,
                         'initial_dim': 64, 'factor': 1, 'output_dim': 64,
                         'cutoff': [200, 800], 'vocab_size': 2000, 'scale_embeds': True},
            'position': {'type': 'sinusoidal_positional', 'init_size': 512, 'embedding_dim': 64,
                         'padding_idx': 1, 'left_pad': False}},
            'embedder_to_indexer_map': {'adaptive': ['roberta'], 'position': ['roberta']},
            'allow_unmatched_keys': True},
        'max_target_positions': 512, 'dropout': 0.1, 'share_decoder_input_output_embed': True,
        'decoder_output_dim': 64, 'decoder_conv_dim': 64, 'decoder_glu': True,
        'decoder_conv_type': 'dynamic', 'weight_softmax': True, 'decoder_attention_heads': 4,
        'weight_dropout': 0.1, 'relu_dropout': 0.0, 'input_dropout': 0.1,
        'decoder_normalize_before': False, 'attention_dropout': 0.1, 'decoder_ffn_embed_dim': 128,
        'decoder_kernel_size_list': [3, 7], 'adaptive_softmax_cutoff': [200, 800],
        'tie_adaptive_weights': True, 'adaptive_softmax_dropout': 0, 'tie_adaptive_proj': False,
        'adaptive_softmax_factor': 1, 'decoder_layers': 2, 'final_norm': False, 'padding_idx': 0,
        'namespace': 'bpe', 'vocab_size': 2000, 'section_attn': False}
    dec = Decoder.from_params(block)
    from tell_b200 import synth
    sd = synth.decoder_state_dict(synth.CFG_TINY, 0)
    missing, unexpected = dec.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    # tied parameters are shared objects, as in the reference
    assert dec.adaptive_softmax.head.word_proj.weight is \
        dec.embedder.token_embedder_adaptive.embeddings[0][0].weight
    n_params = sum(p.numel() for p in dec.parameters())
    ref = {k: v for k, v in sd.items()
           if v.is_floating_point() and 'weights' not in k and 'version' not in k
           and '_float_tensor' not in k and 'word_proj' not in k and 'tail.0.2' not in k
           and 'tail.1.2' not in k}
    assert n_params == sum(v.numel() for v in ref.values())
    bad = dict(block, bogus_key=1)
    with pytest.raises(ConfigurationError):
        Decoder.from_params(bad)


@pytest.mark.parametrize('name,model_cls,decoder_cls,n_ctx', [
    ('9_transformer_objects', 'TransformerFacesObjectModel', 'DynamicConvFacesObjectsDecoder', 4),
    ('8_transformer_faces', 'TransformerFacesModel', 'DynamicConvFacesParallelDecoder', 3),
    ('5_transformer_roberta', 'TransformerFlattenedModel', 'DynamicConvFlattenedDecoder', 2),
    ('4_no_image', 'TransformerFlattenedModel', 'DynamicConvDecoderNoImage', 1)])
def test_registry_builds_the_real_model_blocks(name, model_cls, decoder_cls, n_ctx):
    """The REAL `model:` blocks of expt/nytimes/*/config.yaml (tests/golden/model_blocks.yaml,
    extracted verbatim by oracle/extract_model_blocks.py) construct through the registry unchanged;
    the cfg-2 decoder has the reference's 200,461,312 parameters and strict-loads a
    reference-keyed state dict.  The two frozen encoders are injected small (building
    roberta.large on the CPU adds nothing to what this test checks)."""
    import yaml
    import tell_b200.models as M
    from tell_b200 import synth
    with open(os.path.join(ROOT, 'tests', 'golden', 'model_blocks.yaml')) as f:
        blocks = yaml.safe_load(f)
    resnet = M.ResNetFeatureExtractor((1, 1, 1, 1))
    roberta = M.RobertaEncoder(n_layers=1, embed_dim=64, heads=4, ffn=64, vocab=100, max_pos=20)
    model = M.Model.from_params(blocks[name], resnet=resnet, roberta=roberta)
    assert type(model).__name__ == model_cls
    assert type(model.decoder).__name__ == decoder_cls
    assert len(model.decoder.layers) == 4
    assert len(model.decoder.layers[0].context_attns) == n_ctx
    assert [l.conv.kernel_size for l in model.decoder.layers] == [3, 7, 15, 31]
    n_dec = sum(p.numel() for p in model.decoder.parameters())
    if name == '9_transformer_objects':
        assert n_dec == 200461312
        assert model.weigh_bert and model.bert_weight.numel() == 2     # 1 injected layer + embeddings
        sd = synth.decoder_state_dict(synth.CFG_FULL, 0)
        missing, unexpected = model.decoder.load_state_dict(sd, strict=True)
        assert not missing and not unexpected
    assert all(not p.requires_grad for p in model.resnet.parameters())
    assert all(not p.requires_grad for p in model.roberta.parameters())


def test_allennlp_bridge_is_shipped_and_explains_itself():
    """INTEGRATION.md section 1 tells the maintainer to import tell_b200.allennlp_bridge; neither
    allennlp nor the reference package exists in this image, so the import must fail with the
    reason (not with a missing module)."""
    import importlib
    with pytest.raises(ImportError, match='allennlp'):
        importlib.import_module('tell_b200.allennlp_bridge')


def test_collator_host_packing_and_generations_writer(tmp_path):
    """Host half of the batch-dict producer (no GPU): ragged packing + offsets, and the
    generations.jsonl wire format of tell/commands/evaluate.py:196-215."""
    import json
    import numpy as np
    from tell_b200.data import Collator, write_generations_jsonl
    arrs, off, max_len = Collator.pack_tokens([[0, 5, 2], [0, 2], []])
    assert off.tolist() == [0, 3, 5, 5] and max_len == 3 and [a.tolist() for a in arrs][1] == [0, 2]
    flat, off, max_rows, width = Collator.pack_arrays([np.ones((2, 4)), np.array([[]]), np.zeros((3, 4))])
    assert off.tolist() == [0, 2, 2, 5] and max_rows == 3 and width == 4 and flat[1].size == 0
    flat, off, max_rows, width = Collator.pack_arrays([np.array([[]]), np.array([[]])])
    assert (max_rows, width) == (1, 0)
    with pytest.raises(ValueError):
        Collator.pack_arrays([np.ones((1, 4)), np.ones((1, 5))])
    out = {'captions': ['a cap'], 'generations': ['a gen'], 'copied_texts': ['cp'],
           'metadata': [{'caption': 'A Cap', 'web_url': 'http://x', 'image_path': '/i.jpg', 'context': 'ctx'}]}
    path = str(tmp_path / 'generations.jsonl')
    assert write_generations_jsonl(path, out) == 1
    assert write_generations_jsonl(path, out, nlp=lambda t: {'names': [t.upper()]}) == 1
    assert write_generations_jsonl(path, {'loss': 1.0}) == 0
    rows = [json.loads(l) for l in open(path)]
    assert rows[0] == {'caption': 'a cap', 'raw_caption': 'A Cap', 'generation': 'a gen', 'copied_texts': 'cp',
                       'web_url': 'http://x', 'image_path': '/i.jpg', 'context': 'ctx', 'copied_text': 'cp'}
    assert rows[1]['generated_names'] == ['A GEN'] and rows[1]['context_names'] == ['CTX']
    # Model.forward(evaluate_mode) emits gen_ids, not generations: the writer builds them with the
    # caller's BPE decoder from x[x > 1] (transformer_faces_objects.py:95-96); without one it says so
    model_out = {'captions': ['a cap', 'b cap'], 'metadata': [{'caption': 'A'}, None],
                 'gen_ids': np.array([[0, 17, 23, 2, 1, 1], [0, 9, 2, 1, 1, 1]])}
    with pytest.raises(KeyError, match='decode'):
        write_generations_jsonl(path, model_out)
    assert write_generations_jsonl(path, model_out, decode=lambda ids: ' '.join(str(int(i)) for i in ids)) == 2
    rows = [json.loads(l) for l in open(path)]
    assert rows[-2]['generation'] == '17 23 2' and rows[-1]['generation'] == '9 2'
    assert rows[-1]['raw_caption'] == 'b cap' and rows[-1]['web_url'] is None
