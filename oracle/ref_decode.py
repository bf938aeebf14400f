"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's decode loop on the torch-flavour cell.

Parity status: UNPINNED.  Greedy/beam decoding exist only in the reference's TensorFlow 1.12 graph
(model/components/dynamic_decode.py:17-74, greedy_decoder_cell.py:46-66, beam_search_decoder_cell.py:98-250), which
cannot run here (no TF); the torch flavour has no decode path at all.  This file restates the TF loop
semantics line by line and applies them to the torch cell of ref_model.py.  Convention: decoding starts
from a caller-supplied START id.
"""
import torch
import torch.nn.functional as F

from oracle import ref_model as rm


def cell_step(p, enc, att1, h, c, tok):
    """One step of DecoderWithAttention's loop body (seq2seq_torch.py:309-316) without dropout; returns logits."""
    ctx, _ = rm.attention_forward(p, enc, h, att1)
    gate = torch.sigmoid(F.linear(h, p["f_beta.weight"], p["f_beta.bias"]))
    x = torch.cat([F.embedding(tok, p["embedding.weight"]), gate * ctx], dim=1)
    h, c = rm.lstm_cell(p, x, h, c)
    return F.linear(h, p["fc.weight"], p["fc.bias"]), h, c


def greedy_decode(p, enc, start_id, end_id, max_iter):
    """dynamic_decode.py:38-61 + greedy_decoder_cell.py:53-66.  enc [N,R,C].  Returns ids [N, steps]."""
    N = enc.shape[0]
    att1 = F.linear(enc, p["attention.encoder_att.weight"], p["attention.encoder_att.bias"])
    h, c = rm.init_hidden_state(p, enc)
    tok = torch.full((N,), start_id, dtype=torch.long)
    finished = torch.zeros(N, dtype=torch.bool)
    out = []
    time = 0
    while not bool(finished.all()):                          # condition :38-40
        logits, h, c = cell_step(p, enc, att1, h, c, tok)
        ids = torch.argmax(logits, dim=-1)                   # greedy_decoder_cell.py:58 (lowest index on ties)
        out.append(ids)
        finished = finished | (ids == end_id)                # :62
        finished = finished | torch.tensor(time >= max_iter) # dynamic_decode.py:49-51
        tok = ids
        time += 1
    return torch.stack(out, dim=1)


def add_div_penalty(log_probs, div_gamma, div_prob, u):
    """beam_search_decoder_cell.py:258-287 (Li et al. 2016).  log_probs [N, beam, V]; u: the uniforms [N, beam, V] that
    sample_bernoulli (:253-255, ``tf.greater(p, tf.random_uniform(s))``) would draw — injected so that a test can replay them.
    rank = position in the descending sort of each beam row (tf.nn.top_k(sorted=True): lower index first among equals),
    penalty = log(gamma) * rank where div_prob > u."""
    if div_gamma is None or div_prob is None or div_gamma == 1.0 or div_prob == 0.0:          # :268-273
        return log_probs
    order = torch.argsort(log_probs, dim=-1, descending=True, stable=True)                    # :276  (stable = top_k's tie rule)
    rank = torch.argsort(order, dim=-1)                                                       # :278-280 invert_permutation
    pen = torch.log(torch.tensor(float(div_gamma), dtype=log_probs.dtype)) * rank.to(log_probs.dtype)   # :282
    return log_probs + pen * (div_prob > u).to(log_probs.dtype)                               # :284-287


def beam_decode(p, enc, start_id, end_id, beam, max_iter, finalize="reference", div_gamma=1.0, div_prob=0.0, div_u=None):
    """beam_search_decoder_cell.py:98-250.  Returns ids [N, steps, beam] (time-major inside the cell, batch-major here)
    and the final log-probs [N, beam].  div_u [steps, N*beam, V]: injected uniforms of the diversity penalty."""
    N, R, C = enc.shape
    V = p["fc.weight"].shape[0]
    att1 = F.linear(enc, p["attention.encoder_att.weight"], p["attention.encoder_att.bias"])
    h, c = rm.init_hidden_state(p, enc)
    # tile_beam :332-350
    enc_t = enc.repeat_interleave(beam, dim=0)
    att1_t = att1.repeat_interleave(beam, dim=0)
    h = h.repeat_interleave(beam, dim=0)
    c = c.repeat_interleave(beam, dim=0)
    tok = torch.full((N * beam,), start_id, dtype=torch.long)
    log_probs = torch.zeros(N, beam)                          # :106-107
    finished = torch.zeros(N, beam, dtype=torch.bool)
    ids_t, parents_t = [], []
    time = 0
    fmin = torch.finfo(torch.float32).min
    while not bool(finished.all()):
        logits, h, c = cell_step(p, enc_t, att1_t, h, c, tok)
        step_lp = torch.log_softmax(logits.view(N, beam, V), dim=-1)                      # :146
        one_hot = torch.full((V,), fmin)
        one_hot[end_id] = 0.0
        f = finished.unsqueeze(-1).float()
        step_lp = (1.0 - f) * step_lp + f * one_hot                                       # mask_probs :353-367
        lp = log_probs.unsqueeze(-1) + step_lp                                            # :150
        if div_u is not None:
            lp = add_div_penalty(lp, div_gamma, div_prob, div_u[time].view(N, beam, V))   # :151-152
        flat = lp.reshape(N, beam * V) if time > 0 else lp[:, 0]                          # :156-160
        new_probs, idx = _topk_low_index_first(flat, beam)                                # :161
        new_ids = idx % V                                                                 # :164
        new_parents = idx // V                                                            # :165
        finished = torch.gather(finished, 1, new_parents) | (new_ids == end_id)           # :171-174
        rows = (new_parents + torch.arange(N).unsqueeze(1) * beam).view(-1)               # gather_helper :370-391
        h, c = h[rows], c[rows]
        log_probs = new_probs
        ids_t.append(new_ids)
        parents_t.append(new_parents)
        finished = finished | torch.tensor(time >= max_iter)
        tok = new_ids.view(-1)
        time += 1
    ids = torch.stack(ids_t, dim=1)                # [N, steps, beam]
    parents = torch.stack(parents_t, dim=1)
    if finalize == "reference":
        # finalize :189-250 gathers with the UNCHANGED initial parents range(beam) at every step -> identity
        return ids, log_probs
    return backtrack(ids, parents), log_probs


def backtrack(ids, parents):
    """Lineage-consistent hypotheses (what finalize was presumably meant to do)."""
    N, S, beam = ids.shape
    out = torch.zeros_like(ids)
    cur = torch.arange(beam).unsqueeze(0).repeat(N, 1)
    for t in range(S - 1, -1, -1):
        out[:, t] = torch.gather(ids[:, t], 1, cur)
        cur = torch.gather(parents[:, t], 1, cur)
    return out


def _topk_low_index_first(x, k):
    """tf.nn.top_k: among equal values the lower index comes first."""
    vals, idxs = [], []
    x = x.clone()
    for _ in range(k):
        m = x.max(dim=1).values
        i = (x == m.unsqueeze(1)).float().argmax(dim=1)
        vals.append(m)
        idxs.append(i)
        x[torch.arange(x.shape[0]), i] = float("-inf")
    return torch.stack(vals, 1), torch.stack(idxs, 1)


def truncate_end(ids, end_id):
    """model/evaluation/text.py:95-104."""
    out = []
    for t in ids:
        if t == end_id:
            break
        out.append(int(t))
    return out
