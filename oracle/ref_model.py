"""TEST INFRASTRUCTURE ONLY — CPU restatement (plain PyTorch, fp32/fp64) of the
reference's im2latex training path.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package; the
product (latex_ocr_b200/) never does.

Parity status: PINNED against the reference's own PyTorch modules executed in the
build container (oracle/make_golden.py imports /root/reference unmodified through
oracle/ref_shim.py and writes tests/golden/*.pt; tests/test_oracle_pinned.py
re-checks the restatement against those fixtures, and against the live reference
when /root/reference is present).  The reference itself ships no tests or golden
vectors, so "the reference run here" is the pin.

Every function cites the reference lines it restates
(paths relative to the reference root).

Layout conventions: parameters travel in plain dicts keyed by the reference's
``state_dict`` names (``cnn.0.weight`` ... for the encoder, ``attention.encoder_att.weight``
... for the decoder).
"""
import math

import torch
import torch.nn.functional as F

# conv index in nn.Sequential -> (pool after it or None); model/components/seq2seq_torch.py:31-57
VANILLA_LAYERS = (
    ("cnn.0", 1, (2, 2)),
    ("cnn.3", 1, (2, 2)),
    ("cnn.6", 1, None),
    ("cnn.8", 1, (2, 1)),
    ("cnn.11", 1, (1, 2)),
    ("cnn.14", 0, None),
)


# ----------------------------------------------------------------------------------------------
# encoder
# ----------------------------------------------------------------------------------------------
def timing_signal_nd(channels, height, width, dtype=torch.float32,
                     min_timescale=1.0, max_timescale=1.0e4):
    """seq2seq_torch.py:115-157 (== model/components/positional.py:10-65).
    Returns the additive table [channels, height, width]."""
    num_dims = 2
    nts = channels // (num_dims * 2)
    inc = math.log(float(max_timescale) / float(min_timescale)) / (float(nts) - 1)
    inv = min_timescale * torch.exp(torch.arange(nts).float() * (-inc))
    table = torch.zeros(channels, height, width, dtype=torch.float32)
    for dim, length in enumerate((height, width)):
        pos = torch.arange(length).float()
        st = inv.unsqueeze(1) * pos.unsqueeze(0)                       # [nts, length]
        sig = torch.cat([torch.sin(st), torch.cos(st)], dim=0)          # [2 nts, length]
        lo = dim * 2 * nts
        if dim == 0:
            table[lo:lo + 2 * nts] += sig[:, :, None]
        else:
            table[lo:lo + 2 * nts] += sig[:, None, :]
    return table.to(dtype)


# 'cnn' variant (seq2seq_torch.py:58-86): no asymmetric pools, Conv2d(512,512,(2,4),stride 2,padding 1)+ReLU instead
CNN_LAYERS = (
    ("cnn.0", 1, (2, 2)),
    ("cnn.3", 1, (2, 2)),
    ("cnn.6", 1, None),
    ("cnn.8", 1, None),
    ("cnn.10", 1, None),
    ("cnn.12", 1, None),      # kernel (2,4), stride 2
    ("cnn.14", 0, None),
)


def encoder_forward(p, img, positional=True, keep=None, encoder_cnn="vanilla"):
    """EncoderCNN.forward: 'vanilla' stack seq2seq_torch.py:31-57, 'cnn' stack :58-86, forward :88-100.
    img [N,1,H,W] raw 0..255 floats (img2seq_torch.py:115-117) -> [N,H',W',512]."""
    x = img
    for name, pad, pool in (VANILLA_LAYERS if encoder_cnn == "vanilla" else CNN_LAYERS):
        stride = 2 if (encoder_cnn == "cnn" and name == "cnn.12") else 1
        x = F.relu(F.conv2d(x, p[name + ".weight"], p[name + ".bias"], stride=stride, padding=pad))
        if keep is not None:
            keep[name] = x
        if pool is not None:
            x = F.max_pool2d(x, kernel_size=pool, stride=pool)
    if positional:
        x = x + timing_signal_nd(x.shape[1], x.shape[2], x.shape[3], x.dtype)[None]
    return x.permute(0, 2, 3, 1)


# ----------------------------------------------------------------------------------------------
# decoder (teacher forced), forward exactly as executed by the reference
# ----------------------------------------------------------------------------------------------
def attention_forward(p, enc, h, att1=None, prefix="attention."):
    """Attention.forward seq2seq_torch.py:178-192.  ``att1`` may be passed in
    (hoisted: the reference recomputes it every step at :186)."""
    if att1 is None:
        att1 = F.linear(enc, p[prefix + "encoder_att.weight"], p[prefix + "encoder_att.bias"])
    att2 = F.linear(h, p[prefix + "decoder_att.weight"], p[prefix + "decoder_att.bias"])
    e = F.linear(F.relu(att1 + att2.unsqueeze(1)),
                 p[prefix + "full_att.weight"], p[prefix + "full_att.bias"]).squeeze(2)
    alpha = torch.softmax(e, dim=1)
    ctx = (enc * alpha.unsqueeze(2)).sum(dim=1)
    return ctx, alpha


def init_hidden_state(p, enc):
    """DecoderWithAttention.init_hidden_state seq2seq_torch.py:255-265."""
    m = enc.mean(dim=1)
    return (F.linear(m, p["init_h.weight"], p["init_h.bias"]),
            F.linear(m, p["init_c.weight"], p["init_c.bias"]))


def lstm_cell(p, x, h, c):
    """nn.LSTMCell (seq2seq_torch.py:222, :313): gate order i,f,g,o, two biases."""
    g = F.linear(x, p["decode_step.weight_ih"], p["decode_step.bias_ih"]) + \
        F.linear(h, p["decode_step.weight_hh"], p["decode_step.bias_hh"])
    i, f, gg, o = g.chunk(4, dim=1)
    i, f, gg, o = torch.sigmoid(i), torch.sigmoid(f), torch.tanh(gg), torch.sigmoid(o)
    c2 = f * c + i * gg
    h2 = o * torch.tanh(c2)
    return h2, c2


def decoder_forward(p, encoder_out, encoded_captions, caption_lengths, dropout_mask=None, hoist=False):
    """DecoderWithAttention.forward seq2seq_torch.py:267-320.

    dropout_mask: None (eval / p=0) or a [B, T, D] multiplier tensor (0 or 1/(1-p)) in the
    *sorted* row order, applied to h before ``fc`` (:316).
    hoist=False re-evaluates encoder_att every step like the reference (:186/:309)."""
    B = encoder_out.size(0)
    C = encoder_out.size(-1)
    enc = encoder_out.reshape(B, -1, C)
    R = enc.size(1)
    caption_lengths, sort_ind = caption_lengths.squeeze(1).sort(dim=0, descending=True)
    enc = enc[sort_ind]
    caps = encoded_captions[sort_ind]
    emb = F.embedding(caps, p["embedding.weight"])
    h, c = init_hidden_state(p, enc)
    decode_lengths = (caption_lengths - 1).tolist()
    T = max(decode_lengths)
    V = p["fc.weight"].shape[0]
    preds = torch.zeros(B, T, V, dtype=enc.dtype)
    alphas = torch.zeros(B, T, R, dtype=enc.dtype)
    att1 = None
    if hoist:
        att1 = F.linear(enc, p["attention.encoder_att.weight"], p["attention.encoder_att.bias"])
    for t in range(T):
        bt = sum([l > t for l in decode_lengths])
        ctx, alpha = attention_forward(p, enc[:bt], h[:bt], None if att1 is None else att1[:bt])
        gate = torch.sigmoid(F.linear(h[:bt], p["f_beta.weight"], p["f_beta.bias"]))
        h, c = lstm_cell(p, torch.cat([emb[:bt, t, :], gate * ctx], dim=1), h[:bt], c[:bt])
        hd = h if dropout_mask is None else h * dropout_mask[:bt, t, :]
        preds[:bt, t, :] = F.linear(hd, p["fc.weight"], p["fc.bias"])
        alphas[:bt, t, :] = alpha
    return preds, caps, decode_lengths, alphas, sort_ind


def loss_from_outputs(scores, caps_sorted, decode_lengths, alphas, alpha_c=1.0):
    """img2seq_torch.py:147-159: CE(mean over packed positions) + alpha_c*mean((1-sum_t alpha)^2)."""
    from torch.nn.utils.rnn import pack_padded_sequence
    targets = caps_sorted[:, 1:]
    ps = pack_padded_sequence(scores, decode_lengths, batch_first=True).data
    pt = pack_padded_sequence(targets, decode_lengths, batch_first=True).data
    ce = F.cross_entropy(ps, pt)
    reg = ((1. - alphas.sum(dim=1)) ** 2).mean()
    return ce + alpha_c * reg, ce, reg


def get_loss(p_enc, p_dec, img, formula, dropout_mask=None, hoist=False, positional=True, alpha_c=1.0):
    """img2seq_torch.py:136-159 forward half of getLoss (every row's length := padded length, :144)."""
    imgs = encoder_forward(p_enc, img, positional)
    lengths = torch.LongTensor([[len(i)] for i in formula])
    scores, caps, dl, alphas, sort_ind = decoder_forward(p_dec, imgs, formula, lengths, dropout_mask, hoist)
    loss, ce, reg = loss_from_outputs(scores, caps, dl, alphas, alpha_c)
    return loss, dict(ce=ce, reg=reg, scores=scores, alphas=alphas, enc=imgs)


def adam_step(params, grads, state, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam defaults as used at img2seq_torch.py:86-87 (no weight decay, no amsgrad).
    state: dict(step=int, m={k:tensor}, v={k:tensor}); updates params in place."""
    state["step"] = state.get("step", 0) + 1
    t = state["step"]
    bc1 = 1.0 - b1 ** t
    bc2 = 1.0 - b2 ** t
    for k in params:
        g = grads[k]
        m = state.setdefault("m", {}).setdefault(k, torch.zeros_like(g))
        v = state.setdefault("v", {}).setdefault(k, torch.zeros_like(g))
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        params[k].addcdiv_(m, denom, value=-lr / bc1)


def train_step(p_enc, p_dec, img, formula, opt_state, lr=1e-3, dropout_mask=None, hoist=False,
               positional=True):
    """Full getLoss (img2seq_torch.py:136-172) with autograd on CPU: returns -loss like :172 plus
    the gradients (reference-layout dicts) for parity checks."""
    pe = {k: v.detach().clone().requires_grad_(True) for k, v in p_enc.items()}
    pd = {k: v.detach().clone().requires_grad_(True) for k, v in p_dec.items()}
    loss, aux = get_loss(pe, pd, img, formula, dropout_mask, hoist, positional)
    loss.backward()
    g_enc = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in pe.items()}
    g_dec = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in pd.items()}
    with torch.no_grad():
        adam_step(p_dec, g_dec, opt_state.setdefault("dec", {}), lr)
        adam_step(p_enc, g_enc, opt_state.setdefault("enc", {}), lr)
    return -loss.item(), g_enc, g_dec, aux


# ----------------------------------------------------------------------------------------------
# hand-derived backward of the decoder (the algorithm the CUDA kernels implement; checked against
# autograd of the restatement above in tests/test_oracle_backward.py)
# ----------------------------------------------------------------------------------------------
def decoder_forward_saved(p, enc, caps, T, dropout_mask=None):
    """Hoisted forward that keeps everything the manual backward needs.  enc [B,R,C] (already
    flattened, rows in sorted order), caps [B, T+1] long.  All rows decode T steps (getLoss mode)."""
    B, R, C = enc.shape
    s = dict(enc=enc, caps=caps, T=T)
    s["att1"] = F.linear(enc, p["attention.encoder_att.weight"], p["attention.encoder_att.bias"])
    s["mean"] = enc.mean(dim=1)
    h, c = init_hidden_state(p, enc)
    emb = F.embedding(caps[:, :T], p["embedding.weight"])
    wf = p["attention.full_att.weight"].reshape(-1)
    keys = ("h_prev", "c_prev", "att2", "alpha", "ctx", "gate", "i", "f", "g", "o", "c", "h", "x")
    for k in keys:
        s[k] = []
    for t in range(T):
        att2 = F.linear(h, p["attention.decoder_att.weight"], p["attention.decoder_att.bias"])
        e = (F.relu(s["att1"] + att2[:, None, :]) * wf).sum(-1) + p["attention.full_att.bias"]
        alpha = torch.softmax(e, dim=1)
        ctx = torch.einsum("br,brc->bc", alpha, enc)
        gate = torch.sigmoid(F.linear(h, p["f_beta.weight"], p["f_beta.bias"]))
        x = torch.cat([emb[:, t], gate * ctx], dim=1)
        pre = F.linear(x, p["decode_step.weight_ih"], p["decode_step.bias_ih"]) + \
            F.linear(h, p["decode_step.weight_hh"], p["decode_step.bias_hh"])
        i, f, g, o = pre.chunk(4, dim=1)
        i, f, g, o = torch.sigmoid(i), torch.sigmoid(f), torch.tanh(g), torch.sigmoid(o)
        c2 = f * c + i * g
        h2 = o * torch.tanh(c2)
        for k, v in zip(keys, (h, c, att2, alpha, ctx, gate, i, f, g, o, c2, h2, x)):
            s[k].append(v)
        h, c = h2, c2
    H = torch.stack(s["h"], dim=1)                                        # [B,T,D]
    s["Hd"] = H if dropout_mask is None else H * dropout_mask
    s["dropout_mask"] = dropout_mask
    s["logits"] = F.linear(s["Hd"], p["fc.weight"], p["fc.bias"])          # [B,T,V]
    s["alphas"] = torch.stack(s["alpha"], dim=1)                           # [B,T,R]
    return s


def decoder_backward_manual(p, s, alpha_c=1.0):
    """Backward of loss = CE_mean + alpha_c*mean((1-sum_t alpha)^2) w.r.t. decoder params and enc.
    Mirrors the CUDA schedule: per-step sequential part only touches enc/att1 once (read-only);
    d att1, d enc, and all weight gradients are hoisted out of the time loop."""
    enc, caps, T = s["enc"], s["caps"], s["T"]
    B, R, C = enc.shape
    V = p["fc.weight"].shape[0]
    wf = p["attention.full_att.weight"].reshape(-1)
    g = {k: torch.zeros_like(v) for k, v in p.items()}
    targets = caps[:, 1:T + 1]
    # fused CE fwd/bwd
    lsm = torch.log_softmax(s["logits"], dim=-1)
    ce = -lsm.gather(-1, targets.unsqueeze(-1)).mean()
    dlogits = (lsm.exp() - F.one_hot(targets, V).to(lsm.dtype)) / (B * T)
    S = s["alphas"].sum(dim=1)
    reg = ((1 - S) ** 2).mean()
    dreg = alpha_c * (-2.0) * (1 - S) / (B * R)                           # [B,R], same for every t
    sreg = torch.einsum("btr,br->bt", s["alphas"], dreg)
    g["fc.weight"] = torch.einsum("btv,btd->vd", dlogits, s["Hd"])
    g["fc.bias"] = dlogits.sum(dim=(0, 1))
    dH = dlogits @ p["fc.weight"]                                         # [B,T,D]
    if s["dropout_mask"] is not None:
        dH = dH * s["dropout_mask"]
    Wih, Whh = p["decode_step.weight_ih"], p["decode_step.weight_hh"]
    E = Wih.shape[1] - C
    dh_next = torch.zeros_like(s["h"][0])
    dc_next = torch.zeros_like(dh_next)
    de_all, datt2_all, dctx_all, dgp_all, dG_all = [None] * T, [None] * T, [None] * T, [None] * T, [None] * T
    for t in range(T - 1, -1, -1):
        i, f, gg, o, c, cp = s["i"][t], s["f"][t], s["g"][t], s["o"][t], s["c"][t], s["c_prev"][t]
        tc = torch.tanh(c)
        dh = dH[:, t] + dh_next
        do = dh * tc
        dc = dc_next + dh * o * (1 - tc * tc)
        dG = torch.cat([dc * gg * i * (1 - i), dc * cp * f * (1 - f), dc * i * (1 - gg * gg), do * o * (1 - o)], dim=1)
        dc_next = dc * f
        dx = dG @ Wih
        dgctx = dx[:, E:]
        dh_prev = dG @ Whh
        gate, ctx, alpha = s["gate"][t], s["ctx"][t], s["alpha"][t]
        dctx = dgctx * gate
        dgp = dgctx * ctx * gate * (1 - gate)
        dh_prev = dh_prev + dgp @ p["f_beta.weight"]
        sb = (dctx * ctx).sum(-1) + sreg[:, t]
        dalpha = torch.einsum("bc,brc->br", dctx, enc) + dreg
        de = alpha * (dalpha - sb[:, None])
        mask = (s["att1"] + s["att2"][t][:, None, :]) > 0
        datt2 = wf * torch.einsum("br,bra->ba", de, mask.to(de.dtype))
        dh_prev = dh_prev + datt2 @ p["attention.decoder_att.weight"]
        de_all[t], datt2_all[t], dctx_all[t], dgp_all[t], dG_all[t] = de, datt2, dctx, dgp, dG
        dh_next = dh_prev
    dh0, dc0 = dh_next, dc_next
    de_all = torch.stack(de_all, 1)            # [B,T,R]
    datt2_all = torch.stack(datt2_all, 1)      # [B,T,A]
    dctx_all = torch.stack(dctx_all, 1)
    dgp_all = torch.stack(dgp_all, 1)
    dG_all = torch.stack(dG_all, 1)            # [B,T,4D]
    Hprev = torch.stack(s["h_prev"], 1)
    X = torch.stack(s["x"], 1)
    att2_all = torch.stack(s["att2"], 1)       # [B,T,A]
    # hoisted d att1 / d w_f pass: one sweep over att1
    datt1 = torch.zeros_like(s["att1"])
    dwf = torch.zeros_like(wf)
    for t in range(T):
        pre = s["att1"] + att2_all[:, t][:, None, :]
        datt1 += de_all[:, t][:, :, None] * (pre > 0)
        dwf += torch.einsum("br,bra->a", de_all[:, t], F.relu(pre))
    datt1 = datt1 * wf
    g["attention.full_att.weight"] = dwf.reshape(1, -1)
    g["attention.full_att.bias"] = de_all.sum().reshape(1)
    g["attention.decoder_att.weight"] = torch.einsum("bta,btd->ad", datt2_all, Hprev)
    g["attention.decoder_att.bias"] = datt2_all.sum(dim=(0, 1))
    g["f_beta.weight"] = torch.einsum("btc,btd->cd", dgp_all, Hprev)
    g["f_beta.bias"] = dgp_all.sum(dim=(0, 1))
    g["decode_step.weight_ih"] = torch.einsum("btg,btx->gx", dG_all, X)
    g["decode_step.weight_hh"] = torch.einsum("btg,btd->gd", dG_all, Hprev)
    g["decode_step.bias_ih"] = dG_all.sum(dim=(0, 1))
    g["decode_step.bias_hh"] = dG_all.sum(dim=(0, 1))
    demb = dG_all @ Wih[:, :E]                                           # [B,T,E]
    g["embedding.weight"].index_add_(0, caps[:, :T].reshape(-1), demb.reshape(B * T, -1))
    g["attention.encoder_att.weight"] = torch.einsum("bra,brc->ac", datt1, enc)
    g["attention.encoder_att.bias"] = datt1.sum(dim=(0, 1))
    g["init_h.weight"] = dh0.t() @ s["mean"]
    g["init_h.bias"] = dh0.sum(0)
    g["init_c.weight"] = dc0.t() @ s["mean"]
    g["init_c.bias"] = dc0.sum(0)
    denc = datt1 @ p["attention.encoder_att.weight"]
    denc = denc + torch.einsum("btr,btc->brc", s["alphas"], dctx_all)
    denc = denc + ((dh0 @ p["init_h.weight"] + dc0 @ p["init_c.weight"]) / R)[:, None, :]
    return ce + alpha_c * reg, g, denc


# ----------------------------------------------------------------------------------------------
# synthetic data: white background (255) with 10 % random ink, ragged targets
# ----------------------------------------------------------------------------------------------
def synthetic_batch(B, H, W, V, tmin, tmax, seed=1234, id_pad=None, id_end=None):
    """Images like model/utils/image.py:27-44 (255 padding), targets like
    model/utils/text.py:141-164 (ids, then END, then PAD up to max_len+1)."""
    gen = torch.Generator().manual_seed(seed)
    id_end = V - 1 if id_end is None else id_end          # Vocab appends _UNK,_PAD,_END last: text.py:12-20
    id_pad = V - 2 if id_pad is None else id_pad
    img = torch.full((B, 1, H, W), 255.0)
    ink = torch.rand((B, 1, H, W), generator=gen) < 0.10
    val = torch.randint(0, 255, (B, 1, H, W), generator=gen).float()
    img = torch.where(ink, val, img)
    lens = torch.randint(tmin, tmax + 1, (B,), generator=gen)
    lens[0] = tmax
    L = int(lens.max())
    formula = torch.full((B, L + 1), id_pad, dtype=torch.long)
    for b in range(B):
        n = int(lens[b])
        formula[b, :n] = torch.randint(0, V - 3, (n,), generator=gen)
        formula[b, n] = id_end
    return img, formula


def init_params(V, seed=0, attention_dim=512, embed_dim=512, decoder_dim=512, encoder_dim=512,
                dtype=torch.float32, encoder_cnn="vanilla"):
    """Random-init parameter dicts with the reference's shapes and init rules
    (nn.Conv2d/nn.Linear/nn.LSTMCell defaults; init_weights seq2seq_torch.py:230-236)."""
    gen = torch.Generator().manual_seed(seed)

    def U(shape, bound):
        return ((torch.rand(shape, generator=gen) * 2 - 1) * bound).to(dtype)

    pe = {}
    cin = 1
    convs = ((("cnn.0", 64, (3, 3)), ("cnn.3", 128, (3, 3)), ("cnn.6", 256, (3, 3)), ("cnn.8", 256, (3, 3)), ("cnn.11", 512, (3, 3)),
              ("cnn.14", 512, (3, 3))) if encoder_cnn == "vanilla" else
             (("cnn.0", 64, (3, 3)), ("cnn.3", 128, (3, 3)), ("cnn.6", 256, (3, 3)), ("cnn.8", 256, (3, 3)), ("cnn.10", 512, (3, 3)),
              ("cnn.12", 512, (2, 4)), ("cnn.14", 512, (3, 3))))
    for name, cout, (kh, kw) in convs:
        b = 1.0 / math.sqrt(cin * kh * kw)
        pe[name + ".weight"] = U((cout, cin, kh, kw), b)
        pe[name + ".bias"] = U((cout,), b)
        cin = cout
    pd = {}

    def lin(name, o, i):
        b = 1.0 / math.sqrt(i)
        pd[name + ".weight"] = U((o, i), b)
        pd[name + ".bias"] = U((o,), b)

    lin("attention.encoder_att", attention_dim, encoder_dim)
    lin("attention.decoder_att", attention_dim, decoder_dim)
    lin("attention.full_att", 1, attention_dim)
    pd["embedding.weight"] = U((V, embed_dim), 0.1)
    b = 1.0 / math.sqrt(decoder_dim)
    pd["decode_step.weight_ih"] = U((4 * decoder_dim, embed_dim + encoder_dim), b)
    pd["decode_step.weight_hh"] = U((4 * decoder_dim, decoder_dim), b)
    pd["decode_step.bias_ih"] = U((4 * decoder_dim,), b)
    pd["decode_step.bias_hh"] = U((4 * decoder_dim,), b)
    lin("init_h", decoder_dim, encoder_dim)
    lin("init_c", decoder_dim, encoder_dim)
    lin("f_beta", encoder_dim, decoder_dim)
    pd["fc.weight"] = U((V, decoder_dim), 0.1)
    pd["fc.bias"] = torch.zeros(V, dtype=dtype)
    return pe, pd
