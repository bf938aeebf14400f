"""CPU: the TF-flavour oracle follows attention_cell.py:71-72,87 under dropout — tf.nn.dropout is applied to the LOCAL new_h
(attention + o projection inputs) while AttentionState keeps the UNDROPPED LSTM output as the recurrent h; new_o is dropped
and IS the recurrent o.  Hand-derived expectations with keep_h = 0 (everything dropped) distinguish the two readings: if the
dropped h were the recurrent state, h_1 would be 0."""
import torch

from oracle import ref_tf_model as tfm


def test_recurrent_h_is_undropped_and_attention_sees_dropped_h():
    dims = dict(num_units=8, dim_e=8, dim_o=8, dim_embeddings=8, channels=8)
    p = {k: v.double() for k, v in tfm.init_params_tf(11, seed=3, dims=dims).items()}
    g = torch.Generator().manual_seed(5)
    enc = torch.randn(2, 5, 8, generator=g, dtype=torch.float64)
    att_img = enc @ p["att_img.kernel"]
    c0, h0, o0 = tfm.initial_state(p, enc)
    emb = p["start_token"].expand(2, -1)
    keep_h = torch.zeros(2, 8, dtype=torch.float64)
    keep_o = torch.full((2, 8), 2.0, dtype=torch.float64)
    logits, (c1, h1, o1), alpha = tfm.cell_step(p, enc, att_img, emb, c0, h0, o0, keep_h=keep_h, keep_o=keep_o)
    # hand-derived: LSTMCell output is untouched by the dropout ...
    c_exp, h_exp = tfm.lstm_cell_tf(p, torch.cat([emb, o0], dim=-1), c0, h0)
    assert torch.equal(c1, c_exp) and torch.equal(h1, h_exp) and h1.abs().min() > 0
    # ... while the attention and o saw hd = 0: e = beta . tanh(att_img), o = 2 * tanh(ctx o_W_c)
    a_exp = torch.softmax(torch.tanh(att_img) @ p["att_beta"], dim=1)
    ctx = (a_exp[:, :, None] * enc).sum(dim=1)
    o_exp = torch.tanh(ctx @ p["o_W_c"]) * 2.0
    assert torch.allclose(alpha, a_exp, atol=1e-14) and torch.allclose(o1, o_exp, atol=1e-14)
    assert torch.allclose(logits, o_exp @ p["y_W_o"], atol=1e-14)
    # second step consumes the undropped h_1 and the dropped o_1
    _, (c2, h2, _), _ = tfm.cell_step(p, enc, att_img, emb, c1, h1, o1)
    c2_exp, h2_exp = tfm.lstm_cell_tf(p, torch.cat([emb, o_exp], dim=-1), c_exp, h_exp)
    assert torch.allclose(h2, h2_exp, atol=1e-14) and torch.allclose(c2, c2_exp, atol=1e-14)
