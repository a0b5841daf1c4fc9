"""Batch-dict producer: device-side padding equals the reference's host-side
field padding (right-padded ids with the indexer's padding value, NaN-padded feature arrays), and
the produced dict drives Model.forward."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _instances(rs, B, with_obj=True, empty_faces=()):
    out = []
    for b in range(B):
        nf = 0 if b in empty_faces else rs.randint(1, 5)
        inst = {'context': [0] + list(rs.randint(4, 5000, size=rs.randint(5, 40))) + [2],
                'caption': [0] + list(rs.randint(4, 5000, size=rs.randint(3, 12))) + [2],
                'image': rs.standard_normal((3, 16, 16)).astype(np.float32),
                'face_embeds': rs.standard_normal((nf, 512)).astype(np.float32) if nf else np.array([[]]),
                'metadata': {'caption': 'c%d' % b, 'web_url': 'u', 'image_path': 'p', 'context': 'x'}}
        if with_obj:
            no = rs.randint(0, 7)
            inst['obj_embeds'] = rs.standard_normal((no, 2048)).astype(np.float32) if no else np.array([[]])
        out.append(inst)
    return out


def _ref_pad_tokens(seqs, pad):
    L = max(len(s) for s in seqs)
    return torch.tensor([list(s) + [pad] * (L - len(s)) for s in seqs], dtype=torch.long)


def _ref_pad_arrays(arrays):
    arrs = [np.asarray(a, dtype=np.float32) for a in arrays]
    width = max(a.shape[1] if a.ndim == 2 else 0 for a in arrs)
    rows = max(max(a.shape[0] if a.size else 1 for a in arrs), 1)
    out = np.full((len(arrs), rows, width), np.nan, dtype=np.float32)
    for i, a in enumerate(arrs):
        if a.size:
            out[i, :a.shape[0]] = a
    return torch.from_numpy(out)


def test_collator_matches_host_padding():
    from tell_b200.data import Collator
    rs = np.random.RandomState(0)
    col = Collator('cuda')
    for trial in range(3):            # staging buffers are reused and grown across calls
        inst = _instances(rs, 5 + 3 * trial, empty_faces=(1,))
        batch = col(inst)
        torch.cuda.synchronize()
        assert torch.equal(batch['context']['roberta'].cpu(), _ref_pad_tokens([i['context'] for i in inst], 1))
        assert torch.equal(batch['caption']['roberta'].cpu(), _ref_pad_tokens([i['caption'] for i in inst], 1))
        for key in ('face_embeds', 'obj_embeds'):
            want = _ref_pad_arrays([i[key] for i in inst])
            got = batch[key].cpu()
            assert got.shape == want.shape
            assert torch.equal(torch.isnan(got), torch.isnan(want))
            assert torch.equal(torch.nan_to_num(got), torch.nan_to_num(want))
        assert torch.equal(batch['image'].cpu(), torch.from_numpy(np.stack([i['image'] for i in inst])))
        assert [m['caption'] for m in batch['metadata']] == ['c%d' % b for b in range(len(inst))]
    # nobody has a face: the field collapses to width 0, as ArrayField does for [1,0] arrays
    inst = _instances(rs, 3, with_obj=False, empty_faces=(0, 1, 2))
    batch = col(inst)
    assert batch['face_embeds'].shape == (3, 1, 0) and 'obj_embeds' not in batch
    with pytest.raises(Exception):
        Collator('cpu')


def test_collated_batch_drives_model_forward():
    """Instances -> Collator -> Model.forward: same loss as the batch padded by hand on the host
    (NaN-padded face / object arrays, right-padded ids)."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import test_model_gpu as tm
    from tell_b200 import synth
    from tell_b200.data import Collator
    rs = np.random.RandomState(3)
    B, S, L, P = 3, 14, 5, 2
    cfg = synth.CFG_TINY
    feats = torch.from_numpy(rs.standard_normal((B, P, P, 2048)).astype(np.float32)).bfloat16()
    insts = []
    for b in range(B):
        nf, no = rs.randint(0, 4), rs.randint(1, 5)
        insts.append({'context': [0] + list(rs.randint(4, cfg['vocab'], size=rs.randint(4, S - 2))) + [2],
                      'caption': [0] + list(rs.randint(4, cfg['vocab'], size=rs.randint(3, 8))) + [2],
                      'image': np.zeros((3, 8, 8), dtype=np.float32),
                      'face_embeds': rs.standard_normal((nf, 512)).astype(np.float32) if nf else np.array([[]]),
                      'obj_embeds': rs.standard_normal((no, 2048)).astype(np.float32),
                      'metadata': {}})
    insts[0]['face_embeds'] = rs.standard_normal((2, 512)).astype(np.float32)      # at least one face
    S_pad = max(len(i['context']) for i in insts)
    hid = torch.from_numpy(rs.standard_normal((L, B * S_pad, 1024)).astype(np.float32)).bfloat16()
    _, _, model = tm._tiny_model('bf16x3', tm._StubResNet(feats.cuda()), tm._StubRoberta(hid.cuda(), L - 1))
    model.eval()
    batch = Collator('cuda')(insts)
    out = model(context=batch['context'], image=batch['image'], caption=batch['caption'],
                face_embeds=batch['face_embeds'], obj_embeds=batch['obj_embeds'], metadata=batch['metadata'])
    loss = out['loss'].item()
    ctx_ids = _ref_pad_tokens([i['context'] for i in insts], 1).cuda()
    cap_ids = _ref_pad_tokens([i['caption'] for i in insts], 1).cuda()
    out2 = model(context={'roberta': ctx_ids}, image=torch.zeros(B, 3, 8, 8).cuda(), caption={'roberta': cap_ids},
                 face_embeds=_ref_pad_arrays([i['face_embeds'] for i in insts]).cuda(),
                 obj_embeds=_ref_pad_arrays([i['obj_embeds'] for i in insts]).cuda(), metadata=[{}] * B)
    assert np.isfinite(loss) and abs(loss - out2['loss'].item()) < 1e-6
