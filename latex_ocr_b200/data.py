"""Host-side batching helpers with the reference's semantics (no kernels here):
``minibatches`` model/utils/general.py:15-35, ``pad_batch_images`` model/utils/image.py:27-64
(pad value 255, uint8), ``pad_batch_formulas`` model/utils/text.py:141-164 (+END, PAD to max_len+1)."""
import os

import numpy as np


def minibatches(data_generator, minibatch_size):
    xs, ys = [], []
    for x, y in data_generator:
        if len(xs) == minibatch_size:
            yield xs, ys
            xs, ys = [], []
        xs.append(x)
        ys.append(y)
    if xs:
        yield xs, ys


def pad_batch_images(images, max_shape=None):
    if max_shape is None:
        max_shape = [max(s) for s in zip(*[list(im.shape) for im in images])]
    out = np.full([len(images)] + list(max_shape), 255, dtype=np.uint8)
    for i, im in enumerate(images):
        out[i, :im.shape[0], :im.shape[1]] = im
    return out


def pad_batch_formulas(formulas, id_pad, id_end, max_len=None):
    if max_len is None:
        max_len = max(len(f) for f in formulas)
    out = np.full([len(formulas), max_len + 1], id_pad, dtype=np.int32)
    lengths = np.zeros(len(formulas), dtype=np.int32)
    for i, f in enumerate(formulas):
        out[i, :len(f)] = np.asarray(f, dtype=np.int32)
        out[i, len(f)] = id_end
        lengths[i] = len(f) + 1
    return out, lengths


class SimpleVocab:
    """Duck-type of model/utils/text.py:5-23 ``Vocab`` for synthetic runs: ids 0..n-4 are tokens,
    then _UNK, _PAD, _END (special tokens are appended last, text.py:12-20)."""

    def __init__(self, n_tok):
        self.n_tok = n_tok
        self.id_unk, self.id_pad, self.id_end = n_tok - 3, n_tok - 2, n_tok - 1


# --------------------------------------------------------------------------------------------------------------------
# host input pipeline: greyscale, shape bucketing, the (image, formula) generator, vocabulary
# --------------------------------------------------------------------------------------------------------------------
def get_max_shape(arrays):
    """utils/image.py:16-24."""
    return [max(x) for x in zip(*[list(a.shape) for a in arrays])]


def greyscale(state):
    """utils/image.py:67-71: ITU-R 601 luma of an (H, W, 3) image in float64, truncated to uint8, shape (H, W, 1)."""
    g = state[:, :, 0] * 0.299 + state[:, :, 1] * 0.587 + state[:, :, 2] * 0.114
    return g[:, :, np.newaxis].astype(np.uint8)


def bucket_by_shape(items, bucket_size, shape_of=lambda it: it[0].shape, reference_quirk=False):
    """The ordering rule of DataGenerator.bucket (utils/data_generator.py:84-122) on any iterable: items are appended to a
    per-shape buffer; when an item arrives for a buffer that already holds ``bucket_size`` items, that buffer is emitted first
    (so every emitted run of ``bucket_size`` items has one shape and becomes one same-shape minibatch); the partly filled
    buffers follow at the end in first-seen shape order.  Returns the reordered list.

    ``reference_quirk=True`` reproduces the reference bit for bit: its flush loop (``for (img_path, formula_id) in
    data_buckets[s]``, :107-109) re-binds the names of the item being processed, so after every flush the item appended to the
    emptied buffer is the LAST FLUSHED one again — the arriving item is dropped and its predecessor is listed twice."""
    out, buckets = [], {}
    for it in items:
        s = tuple(shape_of(it))
        buf = buckets.setdefault(s, [])
        if len(buf) == bucket_size:
            out.extend(buf)
            if reference_quirk:
                it = buf[-1]
            del buf[:]
        buf.append(it)
    for buf in buckets.values():
        out.extend(buf)
    return out


def _imread(path):
    """scipy.misc.imread of the reference (utils/data_generator.py:4) is gone from scipy: PIL decodes the same files."""
    from PIL import Image
    return np.asarray(Image.open(path))


class DataGenerator:
    """Drop-in for utils/data_generator.py:36-230: iterates (image, formula) pairs listed in a matching file
    (``<image name> <formula line index>`` per line), with the reference's options — ``img_prepro`` (e.g. ``greyscale``),
    ``form_prepro`` (e.g. ``Vocab.form_prepro``), ``max_iter``, ``max_len`` (longer formulas are skipped), ``iter_mode``
    ('data' -> (img, formula), 'full' -> (img, formula, img_path, formula_id)) and ``bucket`` / ``bucket_size``
    (same-shape images are grouped so that ``minibatches`` yields same-shape batches: no padding waste and one
    workspace / CUDA graph per bucket on the device side)."""

    def __init__(self, path_formulas, dir_images, path_matching, bucket=False, form_prepro=lambda s: s.strip().split(" "),
                 iter_mode="data", img_prepro=lambda x: x, max_iter=None, max_len=None, bucket_size=20, bucket_fix=False):
        from .metrics import load_formulas
        self._dir_images = dir_images
        self._path_matching = path_matching
        self._img_prepro, self._form_prepro = img_prepro, form_prepro
        self._max_iter, self._max_len = max_iter, max_len
        self._iter_mode = iter_mode
        self._length = None
        self._formulas = load_formulas(path_formulas)
        self._listing = None                       # None = stream the matching file; a list after bucketing
        # False (default): the reference's listing bit for bit, including its dropped / duplicated samples (see
        # bucket_by_shape); True: every sample exactly once
        self._bucket_fix = bucket_fix
        if bucket:
            self._listing = self.bucket(bucket_size)

    def _examples(self):
        if self._listing is not None:
            yield from self._listing
            return
        with open(self._path_matching) as f:
            for line in f:
                parts = line.strip().split(" ")
                if len(parts) >= 2:
                    yield parts[0], parts[1]

    def bucket(self, bucket_size):
        """data_generator.py:84-122 (one full pass that decodes every image to learn its shape)."""
        old, self._iter_mode = self._iter_mode, "full"
        full = list(self)
        self._iter_mode = old
        self._length = len(full)
        return [(p, fid) for (_, _, p, fid) in bucket_by_shape(full, bucket_size, reference_quirk=not self._bucket_fix)]

    def _process_instance(self, example):
        img_path, formula_id = example
        img = self._img_prepro(_imread(os.path.join(self._dir_images, img_path)) if not os.path.isabs(img_path) else _imread(img_path))
        try:
            raw = self._formulas[int(formula_id)]
        except KeyError:
            raise KeyError("formula id %s not in the formulas file (%d lines): matching file and formulas disagree"
                           % (formula_id, len(self._formulas)))
        formula = self._form_prepro(raw)
        inst = (img, formula) if self._iter_mode == "data" else (img, formula, img_path, formula_id)
        return inst, (self._max_len is not None and len(formula) > self._max_len)

    def __iter__(self):
        n = 0
        for ex in self._examples():
            if self._max_iter is not None and n >= self._max_iter:
                break
            inst, skip = self._process_instance(ex)
            if skip:
                continue
            n += 1
            yield inst

    def __len__(self):
        if self._length is None:
            self._length = sum(1 for _ in self)
        return self._length


class Vocab:
    """utils/text.py:5-23: tokens of ``config.path_vocab`` (one per line) then the special tokens unk, pad, end."""

    def __init__(self, config):
        self.config = config
        self.tok_to_id = {}
        with open(config.path_vocab) as f:
            for idx, tok in enumerate(f):
                self.tok_to_id[tok.strip()] = idx
        for tok in (config.unk, config.pad, config.end):
            self.tok_to_id[tok] = len(self.tok_to_id)
        self.id_to_tok = {i: t for t, i in self.tok_to_id.items()}
        self.n_tok = len(self.tok_to_id)
        self.id_pad, self.id_end, self.id_unk = (self.tok_to_id[config.pad], self.tok_to_id[config.end], self.tok_to_id[config.unk])

    @property
    def form_prepro(self):
        t2i, unk = self.tok_to_id, self.id_unk
        return lambda formula: [t2i.get(t, unk) for t in formula.strip().split(" ")]


class PinnedBatcher:
    """Pads a minibatch like the reference (pad_batch_images / pad_batch_formulas) straight into PINNED host buffers, reused per
    (batch, shape) bucket, so that the upload is one asynchronous DMA per tensor: uint8 pixels (the normalise / cast to
    float happens inside the conv1 kernel) and int64 token ids.  Returns torch tensors [N,1,H,W] uint8 and [N,L+1] int64."""

    def __init__(self, id_pad, id_end, max_buffers=8):
        self.id_pad, self.id_end = id_pad, id_end
        self._bufs, self._max = {}, max_buffers

    def _buf(self, key, shape, dtype):
        import torch
        b = self._bufs.get(key)
        if b is None:
            if len(self._bufs) >= self._max:
                self._bufs.pop(next(iter(self._bufs)))
            b = torch.empty(shape, dtype=dtype)
            try:
                b = b.pin_memory()
            except RuntimeError:                     # no CUDA driver (CPU-only tests): pageable memory, same values
                pass
            self._bufs[key] = b
        return b

    def images(self, images):
        import torch
        mh, mw = get_max_shape(images)[:2]
        out = self._buf(("img", len(images), mh, mw), (len(images), 1, mh, mw), torch.uint8)
        out.fill_(255)                                                        # utils/image.py:40
        for i, im in enumerate(images):
            a = np.asarray(im)
            a = a[:, :, 0] if a.ndim == 3 else a
            out[i, 0, :a.shape[0], :a.shape[1]] = torch.from_numpy(np.ascontiguousarray(a.astype(np.uint8)))
        return out

    def formulas(self, formulas):
        import torch
        L = max(len(f) for f in formulas)
        out = self._buf(("tok", len(formulas), L), (len(formulas), L + 1), torch.int64)
        out.fill_(self.id_pad)                                                # utils/text.py:157-162
        for i, f in enumerate(formulas):
            if len(f):
                out[i, :len(f)] = torch.as_tensor(list(f), dtype=torch.int64)
            out[i, len(f)] = self.id_end
        return out
