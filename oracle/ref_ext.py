"""TEST INFRASTRUCTURE ONLY — CPU definition (plain PyTorch) of the EXTENSION of BASELINE.json configs[3]: row-encoder biLSTM over
the CNN feature rows + a second decoder layer (latex_ocr_b200/ext.py).  The reference has no such model (SURVEY.md §0:
model/decoder.py:16 only links the im2markup paper), so this is not a restatement of reference code: "parity unpinned — extension";
the LSTM arithmetic is torch.nn.LSTM / nn.LSTMCell itself (gate order i,f,g,o, two bias vectors), the rest re-uses the pinned
restatement oracle/ref_model.py (encoder, attention, loss).

  row encoder : nn.LSTM(512, 256, bidirectional=True, batch_first=True) over each row of the [N,H',W',512] feature map, zero state
  decoder     : layer 1 = the attention LSTM of seq2seq_torch.py:267-320 (unchanged); x_t = dropout(h1_t);
                layer 2 = nn.LSTMCell(D, D) over x_t, zero initial state; logits_t = fc(h2_t)
  loss        : img2seq_torch.py:147-159 (CE over all padded positions + doubly-stochastic regulariser)
"""
import math

import torch
import torch.nn.functional as F

from oracle import ref_model as rm


def init_params_ext(seed=0, channels=512, hidden=256, D=512, dtype=torch.float32):
    gen = torch.Generator().manual_seed(seed)

    def U(shape, bound):
        return ((torch.rand(shape, generator=gen) * 2 - 1) * bound).to(dtype)
    b = 1.0 / math.sqrt(hidden)
    prow = {}
    for suf in ("_l0", "_l0_reverse"):
        prow["lstm.weight_ih" + suf] = U((4 * hidden, channels), b)
        prow["lstm.weight_hh" + suf] = U((4 * hidden, hidden), b)
        prow["lstm.bias_ih" + suf] = U((4 * hidden,), b)
        prow["lstm.bias_hh" + suf] = U((4 * hidden,), b)
    b = 1.0 / math.sqrt(D)
    p2 = {"cell.weight_ih": U((4 * D, D), b), "cell.weight_hh": U((4 * D, D), b), "cell.bias_ih": U((4 * D,), b),
          "cell.bias_hh": U((4 * D,), b)}
    return prow, p2


def lstm_seq(x, w_ih, w_hh, b_ih, b_hh, reverse=False):
    """nn.LSTM single direction, batch_first, zero initial state: x [M,S,I] -> [M,S,H]."""
    M, S, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(M, H)
    c = x.new_zeros(M, H)
    out = [None] * S
    for t in (range(S - 1, -1, -1) if reverse else range(S)):
        g = F.linear(x[:, t], w_ih, b_ih) + F.linear(h, w_hh, b_hh)
        i, f, gg, o = g.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        out[t] = h
    return torch.stack(out, dim=1)


def row_encoder_forward(p, feat):
    """feat [N,H',W',C] -> [N,H',W',2*hidden]: forward || backward halves, exactly nn.LSTM(bidirectional=True) per row."""
    N, Hh, Ww, C = feat.shape
    x = feat.reshape(N * Hh, Ww, C)
    fw = lstm_seq(x, p["lstm.weight_ih_l0"], p["lstm.weight_hh_l0"], p["lstm.bias_ih_l0"], p["lstm.bias_hh_l0"])
    bw = lstm_seq(x, p["lstm.weight_ih_l0_reverse"], p["lstm.weight_hh_l0_reverse"], p["lstm.bias_ih_l0_reverse"],
                  p["lstm.bias_hh_l0_reverse"], reverse=True)
    return torch.cat([fw, bw], dim=2).reshape(N, Hh, Ww, -1)


def decoder2_forward(pd, p2, enc, caps, T, dropout_mask=None):
    """Two-layer teacher-forced decoder; every row decodes T steps (getLoss mode, img2seq_torch.py:144).  enc [B,R,C]."""
    B, R, _ = enc.shape
    emb = F.embedding(caps, pd["embedding.weight"])
    h, c = rm.init_hidden_state(pd, enc)
    D = h.shape[1]
    h2 = enc.new_zeros(B, D)
    c2 = enc.new_zeros(B, D)
    preds, alphas = [], []
    for t in range(T):
        ctx, alpha = rm.attention_forward(pd, enc, h)
        gate = torch.sigmoid(F.linear(h, pd["f_beta.weight"], pd["f_beta.bias"]))
        h, c = rm.lstm_cell(pd, torch.cat([emb[:, t], gate * ctx], dim=1), h, c)
        x = h if dropout_mask is None else h * dropout_mask[:, t]
        g = F.linear(x, p2["cell.weight_ih"], p2["cell.bias_ih"]) + F.linear(h2, p2["cell.weight_hh"], p2["cell.bias_hh"])
        i, f, gg, o = g.chunk(4, dim=1)
        c2 = torch.sigmoid(f) * c2 + torch.sigmoid(i) * torch.tanh(gg)
        h2 = torch.sigmoid(o) * torch.tanh(c2)
        preds.append(F.linear(h2, pd["fc.weight"], pd["fc.bias"]))
        alphas.append(alpha)
    return torch.stack(preds, dim=1), torch.stack(alphas, dim=1)


def get_loss_ext(pe, prow, pd, p2, img, formula, dropout_mask=None):
    feat = rm.encoder_forward(pe, img)
    enc = row_encoder_forward(prow, feat)
    N = enc.shape[0]
    T = formula.shape[1] - 1
    scores, alphas = decoder2_forward(pd, p2, enc.reshape(N, -1, enc.shape[3]), formula, T, dropout_mask)
    loss, ce, reg = rm.loss_from_outputs(scores, formula, [T] * N, alphas)
    return loss, dict(scores=scores, alphas=alphas, enc=enc, feat=feat)


def train_grads_ext(pe, prow, pd, p2, img, formula, dropout_mask=None):
    """Loss and every gradient by autograd (dicts in the layouts of the parameter dicts)."""
    ps = [{k: v.detach().clone().requires_grad_(True) for k, v in d.items()} for d in (pe, prow, pd, p2)]
    loss, aux = get_loss_ext(ps[0], ps[1], ps[2], ps[3], img, formula, dropout_mask)
    loss.backward()
    grads = [{k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in d.items()} for d in ps]
    return loss.item(), grads, aux
