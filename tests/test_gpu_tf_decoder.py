"""-m gpu: the TensorFlow-flavour decoder (Genthial cell) through the C ABI against the CPU restatement
oracle/ref_tf_model.py (parity UNPINNED: the reference runs this flavour only as a TF 1.12 graph, see the oracle's header).
fp32 mode: logits / loss / every gradient to 1e-4 of the tensor's max, token ids exact; bf16 mode: stated looser tolerances."""
import pytest
import torch

from util import Cfg, relerr

pytestmark = pytest.mark.gpu


def _cfg(**kw):
    return Cfg(attn_cell_config={"num_units": 512, "dim_e": 256, "dim_o": 512, "dim_embeddings": 80}, max_length_formula=10, **kw)


def _setup(V=40, N=3, R=21, T=7, seed=3, precision="fp32", impl="simt", **kw):
    from latex_ocr_b200.tf_decoder import Decoder
    from oracle import ref_tf_model as tfm
    g = torch.Generator().manual_seed(seed)
    p = tfm.init_params_tf(V, seed=seed)
    p["y_W_o"] = p["y_W_o"] * 4.0                               # sharper logits: non-degenerate argmax / beams
    enc = torch.relu(torch.randn(N, R, 512, generator=g)) * 0.7
    formula = torch.randint(0, V, (N, T), generator=g)
    dec = Decoder(_cfg(**kw), V, V - 1, device="cuda", precision=precision, impl=impl)
    dec.load_tf_variables(p)
    return tfm, p, dec, enc, formula


def _oracle_grads(tfm, p, enc, formula, lengths, keep_h=None, keep_o=None):
    p = {k: v.clone().double().requires_grad_(True) for k, v in p.items()}
    enc = enc.clone().double().requires_grad_(True)
    kh = None if keep_h is None else keep_h.double()
    ko = None if keep_o is None else keep_o.double()
    logits, alphas = tfm.decoder_train_logits(p, enc, formula, kh, ko)
    loss, ce_words, n_words = tfm.masked_ce(logits, formula, lengths)
    loss.backward()
    return logits.detach(), alphas.detach(), loss.detach(), {k: v.grad for k, v in p.items()}, enc.grad


@pytest.mark.parametrize("dropout", [False, True])
def test_tf_decoder_fp32_matches_oracle(dropout):
    tfm, p, dec, enc, formula = _setup()
    N, T = formula.shape
    lengths = torch.tensor([7, 5, 3])
    kh = ko = None
    if dropout:
        g = torch.Generator().manual_seed(11)
        kh = (torch.rand(N, T, 512, generator=g) < 0.8).float() / 0.8
        ko = (torch.rand(N, T, 512, generator=g) < 0.8).float() / 0.8
    logits, alphas, loss, grads, denc = _oracle_grads(tfm, p, enc, formula, lengths, kh, ko)
    got = dec.train_logits(enc.cuda(), formula.cuda(), kh, ko)
    assert got.shape == logits.shape
    assert relerr(got, logits) < 1e-4
    l, de = dec.loss_and_backward(enc.cuda(), formula.cuda(), lengths, kh, ko)
    torch.cuda.synchronize()
    assert abs(l[0].item() - loss.item()) / abs(loss.item()) < 1e-5
    assert l[3].item() == float(lengths.sum())
    ws = dec._ws[(N, T, enc.shape[1], 1)]
    assert relerr(ws["alphas"], alphas) < 1e-4
    assert relerr(de, denc) < 1e-4
    for k, q in dec.named_parameters():
        assert relerr(q.grad, grads[k].reshape(q.grad.shape)) < 1e-4, k


def test_tf_decoder_bf16_close_to_oracle():
    """bf16 storage / fp32 accumulate, tcgen05 + mma.sync GEMMs, tanh.approx in the attention score: loss within 2e-2 relative,
    gradients within 6e-2 of their max."""
    tfm, p, dec, enc, formula = _setup(N=4, R=40, T=9, precision="bf16", impl="tc")
    lengths = torch.tensor([9, 9, 6, 2])
    logits, alphas, loss, grads, denc = _oracle_grads(tfm, p, enc, formula, lengths)
    l, de = dec.loss_and_backward(enc.cuda(), formula.cuda(), lengths)
    torch.cuda.synchronize()
    assert abs(l[0].item() - loss.item()) / abs(loss.item()) < 2e-2
    assert relerr(de, denc) < 6e-2
    for k, q in dec.named_parameters():
        assert relerr(q.grad, grads[k].reshape(q.grad.shape)) < 6e-2, k


def test_tf_greedy_ids_match_oracle():
    tfm, p, dec, enc, formula = _setup(seed=5, decoding="greedy")
    want = tfm.greedy_decode(p, enc, end_id=dec._id_end, max_iter=dec.max_length_formula + 1)
    out = dec.decode(enc.cuda())
    assert out.ids.shape == want.shape, (out.ids.shape, want.shape)
    assert torch.equal(out.ids.cpu(), want)
    train, test = dec(enc.cuda(), formula.cuda(), 1.0)              # Decoder.__call__ surface: (pred_train, pred_test)
    assert train.shape == (formula.shape[0], formula.shape[1], 40) and torch.equal(test.ids.cpu(), want)


@pytest.mark.parametrize("beam", [2, 5])
def test_tf_beam_ids_match_oracle(beam):
    tfm, p, dec, enc, formula = _setup(seed=7, decoding="beam_search", beam_size=beam)
    want, wlp = tfm.beam_decode(p, enc, end_id=dec._id_end, beam=beam, max_iter=dec.max_length_formula + 1)
    out = dec.decode(enc.cuda())
    assert out.ids.shape == want.shape, (out.ids.shape, want.shape)
    assert torch.equal(out.ids.cpu(), want)
