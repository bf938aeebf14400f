"""TEST INFRASTRUCTURE ONLY — loader for the UNMODIFIED reference modules.

Imports ``model/components/seq2seq_torch.py`` of LinXueyuanStdio/LaTeX_OCR from
``/root/reference`` (this container only; the GPU box has no reference tree) so
that ``oracle/make_golden.py`` and the CPU tests can pin ``oracle/ref_model.py``
(our restatement) against the reference's own PyTorch code.

Three arithmetic-neutral shims:
  * ``import tensorflow`` at seq2seq_torch.py:9 is unused -> stub module with a
    ``__spec__`` (torchvision/dynamo call ``find_spec`` on it).
  * ``add_timing_signal_nd_torch`` adds in place on a ReLU output
    (seq2seq_torch.py:156) which breaks autograd -> we call it on a clone.
  * ``pack_padded_sequence`` unpacking (img2seq_torch.py:151-152) is restated in
    ``ref_get_loss`` with ``.data``.

Nothing in the product package imports this file.
"""
import importlib
import importlib.machinery
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_reference_root() -> str:
    """Search order (SURVEY.md §8-c "GPU-box note"): $LATEX_OCR_REFERENCE, /root/reference (the build container),
    then <repo>/baseline/_ref (where a driver-side install of the reference lands on the GPU box; git-ignored)."""
    cands = [os.environ.get("LATEX_OCR_REFERENCE"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "model", "components", "seq2seq_torch.py")):
            return c
    return cands[1]


REFERENCE_ROOT = _find_reference_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "components", "seq2seq_torch.py"))


def _stub(name: str) -> None:
    if name in sys.modules:
        return
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    sys.modules[name] = m


def load_reference_components():
    """Returns the reference module ``model.components.seq2seq_torch``."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _stub("tensorflow")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    mod = importlib.import_module("model.components.seq2seq_torch")
    import torch
    mod.device = torch.device("cpu")
    return mod


class RefConfig:
    """Duck-typed stand-in for model/utils/general.py:88 ``Config``."""

    def __init__(self, encoder_cnn="vanilla", positional_embeddings=True):
        self.encoder_cnn = encoder_cnn
        self.positional_embeddings = positional_embeddings


def build_reference_models(vocab_size, encoder_cnn="vanilla", positional_embeddings=True,
                           attention_dim=512, embed_dim=512, decoder_dim=512, dropout=0.5):
    """EncoderCNN + DecoderWithAttention exactly as img2seq_torch.py:77-82 builds them."""
    mod = load_reference_components()
    enc = mod.EncoderCNN(RefConfig(encoder_cnn, positional_embeddings))
    # Shim 1: out-of-place timing signal (forward values identical).
    orig = enc.add_timing_signal_nd_torch
    enc.add_timing_signal_nd_torch = lambda x, *a, **k: orig(x.clone(), *a, **k)
    dec = mod.DecoderWithAttention(attention_dim=attention_dim, embed_dim=embed_dim,
                                   decoder_dim=decoder_dim, vocab_size=vocab_size, dropout=dropout)
    return enc, dec


def ref_get_loss(enc, dec, img, formula, alpha_c=1.0):
    """img2seq_torch.py:136-159 (forward + loss) on the reference modules; returns
    (loss tensor with graph, scores, alphas)."""
    import torch
    from torch.nn.utils.rnn import pack_padded_sequence
    imgs = enc(img)
    scores, caps_sorted, decode_lengths, alphas, sort_ind = dec(
        imgs, formula, torch.LongTensor([[len(i)] for i in formula]))
    targets = caps_sorted[:, 1:]
    packed_scores = pack_padded_sequence(scores, decode_lengths, batch_first=True).data   # Shim 2
    packed_targets = pack_padded_sequence(targets, decode_lengths, batch_first=True).data
    loss = torch.nn.functional.cross_entropy(packed_scores, packed_targets)
    loss = loss + alpha_c * ((1. - alphas.sum(dim=1)) ** 2).mean()
    return loss, scores, alphas
