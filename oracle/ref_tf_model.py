"""TEST INFRASTRUCTURE ONLY — CPU restatement (plain PyTorch) of the reference's TensorFlow-flavour decoder
(SURVEY.md §8-a row a7, §8-A.2): the Genthial attention cell, its teacher-forced training pass, the masked
cross-entropy, and greedy / beam decoding with that cell.

Parity status: UNPINNED.  The reference executes this flavour only as a TensorFlow 1.12 graph (absent here and not
installable) and ships no test, checkpoint or golden output for it.  Pieces that live inside TensorFlow and not in
the reference repository are restated from their published definitions:
  * tf.contrib.rnn.LSTMCell (tensorflow==1.12.2, requirements.txt:1; call sites model/decoder.py:54,62): one kernel
    [input+num_units, 4*num_units], one bias, gate order i, j, f, o, c' = sigmoid(f + forget_bias)*c + sigmoid(i)*tanh(j),
    h' = sigmoid(o)*tanh(c'), forget_bias = 1.0; state tuple (c, h).
  * tf.layers.dense / tf.get_variable default initialiser: glorot_uniform; biases zero.
Every function cites the reference lines it follows (paths relative to the reference root).

Parameters travel in a plain dict keyed by the TF variable names (scope prefixes dropped):
  embedding_table [V,E]  start_token [E]                         model/decoder.py:42-46
  att_img.kernel [C,A]                                           components/attention_mechanism.py:43
  att_h.kernel [D,A]  att_beta [A]                               :79, :86
  W_c_0 [C,D] b_c_0 [D]  W_h_0 [C,D] b_h_0 [D]  W_o_0 [C,O] b_o_0 [O]    :145-153, attention_cell.py:51-56
  lstm.kernel [E+O+D, 4D]  lstm.bias [4D]                        decoder.py:54 (TF LSTMCell)
  o_W_c [C,O]  o_W_h [D,O]  y_W_o [O,V]                          attention_cell.py:76-78
"""
import math

import torch
import torch.nn.functional as F

from oracle import ref_decode as rd

DIMS = dict(num_units=512, dim_e=256, dim_o=512, dim_embeddings=80, channels=512)     # configs/model.json:8-11


def init_params_tf(V, seed=0, dims=None, dtype=torch.float32):
    """Random init with the reference's rules: glorot-uniform variables (tf.get_variable / tf.layers.dense defaults), zero LSTM bias, embedding rows U(-1,1) then
    L2-normalised (decoder.py:98-105)."""
    d = dict(DIMS, **(dims or {}))
    D, A, O, E, C = d["num_units"], d["dim_e"], d["dim_o"], d["dim_embeddings"], d["channels"]
    gen = torch.Generator().manual_seed(seed)

    def glorot(fan_in, fan_out, shape=None):
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        return ((torch.rand(shape or (fan_in, fan_out), generator=gen) * 2 - 1) * lim).to(dtype)

    def emb(shape):
        t = torch.rand(shape, generator=gen) * 2 - 1
        return F.normalize(t, dim=-1).to(dtype)

    p = {"embedding_table": emb((V, E)), "start_token": emb((E,)),
         "att_img.kernel": glorot(C, A), "att_h.kernel": glorot(D, A), "att_beta": glorot(A, 1).reshape(A),
         "lstm.kernel": glorot(E + O + D, 4 * D), "lstm.bias": torch.zeros(4 * D, dtype=dtype),
         "o_W_c": glorot(C, O), "o_W_h": glorot(D, O), "y_W_o": glorot(O, V)}
    for n, dim in (("c", D), ("h", D), ("o", O)):
        p["W_%s_0" % n] = glorot(C, dim)
        # tf.get_variable without initializer -> glorot_uniform also for rank-1 shapes (fan_in = fan_out = dim)
        p["b_%s_0" % n] = glorot(dim, dim, shape=(dim,))
    return p


def initial_state(p, enc):
    """AttentionMechanism.initial_state (attention_mechanism.py:145-153) for c, h (LSTMStateTuple order) and o
    (attention_cell.py:51-56): tanh(mean_r(img) @ W + b)."""
    m = enc.mean(dim=1)
    return tuple(torch.tanh(m @ p["W_%s_0" % n] + p["b_%s_0" % n]) for n in ("c", "h", "o"))


def lstm_cell_tf(p, x, c, h):
    """tf.contrib.rnn.LSTMCell.call (TF 1.12): gates i, j, f, o from one matmul over [x ; h]; forget_bias 1.0."""
    D = h.shape[1]
    z = torch.cat([x, h], dim=1) @ p["lstm.kernel"] + p["lstm.bias"]
    i, j, f, o = z[:, :D], z[:, D:2 * D], z[:, 2 * D:3 * D], z[:, 3 * D:]
    c2 = torch.sigmoid(f + 1.0) * c + torch.sigmoid(i) * torch.tanh(j)
    h2 = torch.sigmoid(o) * torch.tanh(c2)
    return c2, h2


def attention_context(p, enc, att_img, h):
    """AttentionMechanism.context / compute_attention (attention_mechanism.py:50-94): no biases, tanh."""
    att_h = h @ p["att_h.kernel"]
    e = torch.tanh(att_img + att_h[:, None, :]) @ p["att_beta"]
    a = torch.softmax(e, dim=1)
    return (a[:, :, None] * enc).sum(dim=1), a


def cell_step(p, enc, att_img, emb, c, h, o, keep_h=None, keep_o=None):
    """AttentionCell.step (attention_cell.py:58-89).  keep_h / keep_o: optional dropout multipliers (already scaled by
    1/keep_prob like tf.nn.dropout) for new_h and new_o.  Returns logits, (c, h, o), alpha."""
    x = torch.cat([emb, o], dim=-1)                                   # :70
    c2, h2 = lstm_cell_tf(p, x, c, h)                                 # :71  new_h, new_cell_state = cell(x, prev)
    hd = h2 if keep_h is None else h2 * keep_h                        # :72  dropout on the LOCAL new_h only
    ctx, a = attention_context(p, enc, att_img, hd)                   # :75
    o2 = torch.tanh(hd @ p["o_W_h"] + ctx @ p["o_W_c"])               # :82
    if keep_o is not None:
        o2 = o2 * keep_o                                              # :83
    # :87 new_state = AttentionState(new_cell_state, new_o): the recurrent h is the UNDROPPED LSTM output, the recurrent o
    # is the dropped new_o
    return o2 @ p["y_W_o"], (c2, h2, o2), a                           # :84-87


def decoder_train_logits(p, enc, formula, keep_h=None, keep_o=None):
    """Decoder.__call__ training branch (decoder.py:48-57): inputs = [start_token ; E[formula[:, :-1]]] (get_embeddings
    :75-96), tf.nn.dynamic_rnn over all T columns.  enc [N,R,C], formula [N,T] -> logits [N,T,V], alphas [N,T,R]."""
    N, T = formula.shape
    att_img = enc @ p["att_img.kernel"]                               # attention_mechanism.py:43 (once)
    c, h, o = initial_state(p, enc)
    emb_all = torch.cat([p["start_token"].expand(N, 1, -1), F.embedding(formula[:, :-1], p["embedding_table"])], dim=1)
    logits, alphas = [], []
    for t in range(T):
        lg, (c, h, o), a = cell_step(p, enc, att_img, emb_all[:, t], c, h, o,
                                     None if keep_h is None else keep_h[:, t], None if keep_o is None else keep_o[:, t])
        logits.append(lg)
        alphas.append(a)
    return torch.stack(logits, dim=1), torch.stack(alphas, dim=1)


def masked_ce(logits, formula, lengths):
    """img2seq.py:68-71: sparse softmax CE, boolean_mask(sequence_mask(formula_length)), mean.  lengths include END
    (utils/text.py:157-162).  Returns (loss, ce_words, n_words) (img2seq.py:74-75)."""
    N, T, V = logits.shape
    ce = F.cross_entropy(logits.reshape(N * T, V), formula.reshape(N * T), reduction="none").reshape(N, T)
    mask = torch.arange(T)[None, :] < lengths[:, None]
    sel = ce[mask]
    return sel.mean(), sel.sum(), lengths.sum()


def greedy_decode(p, enc, end_id, max_iter):
    """dynamic_decode.py:38-61 + greedy_decoder_cell.py:38-66 on the Genthial cell (start = learned start_token)."""
    N = enc.shape[0]
    att_img = enc @ p["att_img.kernel"]
    c, h, o = initial_state(p, enc)
    emb = p["start_token"].expand(N, -1)
    finished = torch.zeros(N, dtype=torch.bool)
    out, time = [], 0
    while not bool(finished.all()):
        logits, (c, h, o), _ = cell_step(p, enc, att_img, emb, c, h, o)
        ids = torch.argmax(logits, dim=-1)
        out.append(ids)
        finished = finished | (ids == end_id) | torch.tensor(time >= max_iter)
        emb = F.embedding(ids, p["embedding_table"])
        time += 1
    return torch.stack(out, dim=1)


def beam_decode(p, enc, end_id, beam, max_iter, div_gamma=1.0, div_prob=0.0, div_u=None):
    """beam_search_decoder_cell.py:98-187 on the Genthial cell (diversity penalty off, model.json:15-16); finalize is the
    reference's identity gather (:189-250).  Returns ids [N, steps, beam], log-probs [N, beam]."""
    N, R, C = enc.shape
    V = p["y_W_o"].shape[1]
    att_img = enc @ p["att_img.kernel"]
    c, h, o = initial_state(p, enc)
    rep = lambda x: x.repeat_interleave(beam, dim=0)                  # tile_beam :332-350
    enc_t, att_t, c, h, o = rep(enc), rep(att_img), rep(c), rep(h), rep(o)
    emb = p["start_token"].expand(N * beam, -1)
    log_probs = torch.zeros(N, beam)
    finished = torch.zeros(N, beam, dtype=torch.bool)
    ids_t, time = [], 0
    fmin = torch.finfo(torch.float32).min
    while not bool(finished.all()):                                                  # dynamic_decode.py:38-40
        logits, (c, h, o), _ = cell_step(p, enc_t, att_t, emb, c, h, o)
        lp = F.log_softmax(logits.reshape(N, beam, V), dim=-1)                       # :146
        fin_row = torch.full((V,), fmin)
        fin_row[end_id] = 0.0
        f = finished[:, :, None].float()
        lp = (1.0 - f) * lp + f * fin_row                                            # mask_probs :353-367
        total = log_probs[:, :, None] + lp                                           # :150
        if div_u is not None:
            total = rd.add_div_penalty(total, div_gamma, div_prob, div_u[time].view(N, beam, V))    # :151-152
        flat = total[:, 0] if time == 0 else total.reshape(N, beam * V)              # :156-160
        vals, idx = rd._topk_low_index_first(flat, beam)                             # :161
        ids, parents = idx % V, idx // V                                             # :164-165
        finished = torch.gather(finished, 1, parents) | (ids == end_id)              # :171-174
        g = (parents + torch.arange(N)[:, None] * beam).reshape(-1)                  # gather_helper :370-391
        c, h, o = c[g], h[g], o[g]
        log_probs = vals
        ids_t.append(ids)
        finished = finished | torch.tensor(time >= max_iter)                         # dynamic_decode.py:49-51
        emb = F.embedding(ids.reshape(-1), p["embedding_table"])
        time += 1
    return torch.stack(ids_t, dim=1), log_probs
