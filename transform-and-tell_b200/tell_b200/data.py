"""Batch-dict producer and generations writer: the step either side of the
hot path.

`Collator` turns a list of instances (token id lists, a preprocessed image, ragged face / object
feature arrays, metadata) into exactly the batch dict `Model.forward` takes
(tell/models/transformer_faces_objects.py:67-75).  The reference builds it field by field on the
host -- allennlp `TextField.as_tensor` (pads with the indexer's padding value on the right,
tell/data/token_indexers/roberta_indexer.py:185-200), `ArrayField(padding_value=nan)`
(tell/data/dataset_readers/nytimes_faces_ner_matched.py:213-217), `ImageField.as_tensor`
(tell/data/fields/image_field.py:42-44) -- and moves each tensor to the GPU separately
(tell/training/callback_apex_trainer.py:193).  Here every ragged field is packed into ONE pinned
staging buffer with its row offsets, uploaded with ONE asynchronous copy per dtype, and padded on
the device by `tt_pad_ragged_*`; images go through one pinned [B,3,H,W] buffer.

`write_generations_jsonl` is the wire format of `tell/commands/evaluate.py:179-223`
(`generations.jsonl`); the spaCy-derived keys of that file are produced only when an `nlp`
callable is supplied (spaCy is a host-side dependency outside this path)."""
import json

import numpy as np
import torch

from . import _lib
from ._lib import c_float, c_int, c_ll, c_void_p


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


class _Staging:
    """Growable pinned host buffer + device twin for one dtype."""

    def __init__(self, dtype, device):
        self.dtype, self.device = dtype, device
        self.host = torch.empty(0, dtype=dtype).pin_memory()
        self.dev = torch.empty(0, dtype=dtype, device=device)
        self.copied = None      # event: the last asynchronous copy has read the pinned buffer

    def upload(self, chunks):
        """chunks: list of 1-D numpy arrays -> device view of their concatenation (async copy)."""
        if self.copied is not None:
            self.copied.synchronize()     # the DMA of the previous batch must be done with `host`
        n = sum(int(c.size) for c in chunks)
        if n > self.host.numel():
            cap = max(n, 2 * self.host.numel(), 1024)
            self.host = torch.empty(cap, dtype=self.dtype).pin_memory()
            self.dev = torch.empty(cap, dtype=self.dtype, device=self.device)
        h = self.host.numpy()
        o = 0
        for c in chunks:
            h[o:o + c.size] = c
            o += c.size
        self.dev[:n].copy_(self.host[:n], non_blocking=True)
        self.copied = torch.cuda.Event()
        self.copied.record()
        return self.dev[:n]


class Collator:
    """instances -> Model.forward batch dict on `device`.

    An instance is a dict with
      'context', 'caption' : sequences of token ids (already BPE-indexed, <s> ... </s>)
      'image'              : float array [3,H,W] (after the reader's torchvision preprocess)
      'face_embeds'        : float array [n_faces, 512]  (or shape [1,0] / empty when there are none)
      'obj_embeds'         : float array [n_objects, 2048] (optional)
      'metadata'           : anything (passed through as a list)
    """

    def __init__(self, device='cuda', index='roberta', padding_value=1):
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise _lib.TtError('Collator pads on the device; no CPU fallback')
        self.index, self.pad = index, int(padding_value)
        self._stage = {}           # one staging buffer per (field, dtype): no sharing inside a batch
        self._img, self._img_copied = None, None

    # -- host-side packing (pure numpy, testable without a GPU)
    @staticmethod
    def pack_tokens(seqs):
        arrs = [np.asarray(s, dtype=np.int64).reshape(-1) for s in seqs]
        off = np.zeros(len(arrs) + 1, dtype=np.int64)
        np.cumsum([a.size for a in arrs], out=off[1:])
        return arrs, off, max([a.size for a in arrs] + [0])

    @staticmethod
    def pack_arrays(arrays):
        """Ragged [n_i, width] arrays (allennlp ArrayField semantics: empty instances are [1,0])."""
        arrs = [np.asarray(a, dtype=np.float32) for a in arrays]
        arrs = [a.reshape(0, 0) if a.size == 0 else a.reshape(a.shape[0], -1) for a in arrs]
        width = max([a.shape[1] for a in arrs] + [0])
        for a in arrs:
            if a.size and a.shape[1] != width:
                raise ValueError('ragged arrays of one field must share their feature width')
        rows = [a.shape[0] for a in arrs]
        off = np.zeros(len(arrs) + 1, dtype=np.int64)
        np.cumsum(rows, out=off[1:])
        # ArrayField pads to the longest instance; an instance without rows still occupies one
        # (all-padding) row, as its [1, 0] array does in the reference
        return [a.reshape(-1) for a in arrs], off, max(rows + [1]), width

    # -- device-side padding
    def _staging(self, field, dtype):
        key = (field, dtype)
        if key not in self._stage:
            self._stage[key] = _Staging(dtype, self.device)
        return self._stage[key]

    def _tokens(self, field, seqs, fill):
        arrs, off, max_len = self.pack_tokens(seqs)
        B = len(seqs)
        blob = self._staging(field, torch.int64).upload(arrs + [off])
        n = int(off[-1])
        out = torch.empty((B, max_len), dtype=torch.int64, device=self.device)
        _lib.call('tt_pad_ragged_i64', c_void_p(blob.data_ptr()), c_void_p(blob[n:].data_ptr()),
                  c_void_p(out.data_ptr()), c_int(B), c_int(max_len), c_ll(fill), _stream())
        return out

    def _arrays(self, field, arrays):
        flat, off, max_rows, width = self.pack_arrays(arrays)
        B = len(arrays)
        out = torch.empty((B, max_rows, width), dtype=torch.float32, device=self.device)
        if width == 0:
            return out
        blob = self._staging(field, torch.float32).upload(flat)
        offs = self._staging(field, torch.int64).upload([off])
        _lib.call('tt_pad_ragged_f32', c_void_p(blob.data_ptr()), c_void_p(offs.data_ptr()),
                  c_void_p(out.data_ptr()), c_int(B), c_int(max_rows), c_int(width),
                  c_float(float('nan')), _stream())
        return out

    def __call__(self, instances):
        B = len(instances)
        batch = {'context': {self.index: self._tokens('context', [i['context'] for i in instances], self.pad)},
                 'caption': {self.index: self._tokens('caption', [i['caption'] for i in instances], self.pad)}}
        imgs = [np.asarray(i['image'], dtype=np.float32) for i in instances]
        shape = (B,) + imgs[0].shape
        if self._img_copied is not None:
            self._img_copied.synchronize()
        if self._img is None or tuple(self._img.shape) != shape:
            self._img = torch.empty(shape, dtype=torch.float32).pin_memory()
        for k, im in enumerate(imgs):
            self._img[k].copy_(torch.from_numpy(im))
        batch['image'] = self._img.to(self.device, non_blocking=True)
        self._img_copied = torch.cuda.Event()
        self._img_copied.record()
        batch['face_embeds'] = self._arrays('face_embeds', [i['face_embeds'] for i in instances])
        if any('obj_embeds' in i for i in instances):
            batch['obj_embeds'] = self._arrays('obj_embeds',
                                               [i.get('obj_embeds', np.zeros((1, 0))) for i in instances])
        batch['metadata'] = [i.get('metadata') for i in instances]
        return batch


def write_generations_jsonl(path, output_dict, nlp=None, extra=None, decode=None):
    """Append one JSON line per sample, keys as tell/commands/evaluate.py:196-215.  `nlp`: optional
    callable text -> dict of the spaCy-derived keys for that text (names / entities / readability);
    without it those keys are omitted.  `decode`: callable ids (1-D int array with <s> and <pad>
    removed, i.e. x[x > 1] as in transformer_faces_objects.py:96) -> text, used when output_dict
    carries `gen_ids` (what Model.forward emits in evaluate_mode) but no `generations`: the reference
    decodes them with the RoBERTa BPE inside forward, host-side string work this package leaves to
    the caller."""
    if 'captions' not in output_dict or output_dict['captions'] is None:
        return 0
    captions = output_dict['captions']
    generations = output_dict.get('generations')
    if generations is None:
        if decode is None or 'gen_ids' not in output_dict:
            raise KeyError("output_dict has no 'generations': pass decode= (ids -> text) to build them "
                           "from 'gen_ids'")
        generations = []
        for row in np.asarray(output_dict['gen_ids']):
            generations.append(decode(row[row > 1]))        # "We ignore <s> and <pad>" (:95-96)
    metadatas = output_dict['metadata']
    copied = output_dict.get('copied_texts', ['' for _ in captions])
    n = 0
    with open(path, 'a') as f:
        for i, caption in enumerate(captions):
            m = metadatas[i] or {}
            obj = {'caption': caption, 'raw_caption': m.get('caption', caption), 'generation': generations[i],
                   'copied_texts': copied[i], 'web_url': m.get('web_url'), 'image_path': m.get('image_path'),
                   'context': m.get('context')}
            if nlp is not None:
                for prefix, text in (('caption', m.get('caption', caption)), ('generated', generations[i]),
                                     ('context', m.get('context') or '')):
                    for k, v in nlp(text).items():
                        obj['%s_%s' % (prefix, k)] = v
            if 'copied_texts' in output_dict:
                obj['copied_text'] = output_dict['copied_texts'][i]
            if extra:
                obj.update(extra(i))
            f.write(json.dumps(obj) + '\n')
            n += 1
    return n
