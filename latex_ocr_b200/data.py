"""Host-side batching helpers with the reference's semantics (no kernels here):
``minibatches`` model/utils/general.py:15-35, ``pad_batch_images`` model/utils/image.py:27-64
(pad value 255, uint8), ``pad_batch_formulas`` model/utils/text.py:141-164 (+END, PAD to max_len+1)."""
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
