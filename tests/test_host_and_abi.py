"""CPU (-m "not gpu"): host-side logic and the C-ABI surface (library loads, exports every declared symbol,
struct mirror matches, argument validation happens before any GPU work)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from latex_ocr_b200 import _lib, data


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    for name in _lib.declared_symbols():
        assert hasattr(L, name), name
    assert L.lo_version() >= 100
    assert L.lo_sizeof_decoder_args() == ctypes.sizeof(_lib.DecoderArgs)
    out = subprocess.run(["nm", "-D", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for name in _lib.declared_symbols():
        assert (" T " + name) in out, name


def test_argument_validation_without_gpu():
    L = _lib.lib()
    assert L.lo_gemm(None, 0, None, 0, None, 0, 4, 4, 4, 4, 1, 1, 4, 4, 1, 0, 0, 0, None, 0, 0, 0, None) == -1
    assert b"null pointer" in L.lo_last_error()
    assert L.lo_attention_workspace_bytes(64, 512) == 4096 + 64 * 16 * 514 * 4
    assert L.lo_decoder_forward(None, 0, None) == -1
    from latex_ocr_b200 import tf_decoder
    L = tf_decoder._bind()
    assert L.lo_sizeof_tfdec_args() == ctypes.sizeof(tf_decoder.TfDecArgs)
    assert L.lo_tfdec_forward(None, 0, None) == -1
    a = tf_decoder.TfDecArgs()
    a.B, a.T, a.R, a.C, a.A, a.D, a.O, a.E, a.V, a.dt = 64, 150, 868, 512, 256, 512, 512, 80, 500, 1
    assert L.lo_tfdec_workspace_bytes(ctypes.byref(a)) > 0
    a.C = 300                                            # unsupported channel count is refused before any GPU work
    assert L.lo_tfdec_forward(ctypes.byref(a), 0, None) == -1 and b"channels" in L.lo_last_error()


def test_cpu_tensors_are_refused():
    with pytest.raises(_lib.LatexOcrB200Error):
        _lib.ptr(torch.zeros(3))
    from latex_ocr_b200.img2seq import Img2SeqModel
    from util import Cfg
    with pytest.raises(_lib.LatexOcrB200Error):
        Img2SeqModel(Cfg(), n_tok=10, device="cpu")


def test_batching_helpers_follow_reference_rules():
    imgs = [np.zeros((3, 5, 1), np.uint8), np.ones((4, 2, 1), np.uint8)]
    b = data.pad_batch_images(imgs)
    assert b.shape == (2, 4, 5, 1) and b.dtype == np.uint8
    assert b[0, 3, 0, 0] == 255 and b[1, 0, 4, 0] == 255 and b[1, 3, 1, 0] == 1
    f, l = data.pad_batch_formulas([[1, 2, 3], [4]], id_pad=8, id_end=9)
    assert f.tolist() == [[1, 2, 3, 9], [4, 9, 8, 8]] and l.tolist() == [4, 2]
    got = list(data.minibatches(((i, -i) for i in range(5)), 2))
    assert got == [([0, 1], [0, -1]), ([2, 3], [-2, -3]), ([4], [-4])]
    assert list(data.minibatches(iter(()), 3)) == []
    v = data.SimpleVocab(10)
    assert (v.id_unk, v.id_pad, v.id_end) == (7, 8, 9)


def test_flat_store_layout_cpu():
    from latex_ocr_b200.decoder import decoder_specs
    from latex_ocr_b200.params import FlatStore
    S = FlatStore(decoder_specs(512, 512, 512, 500, 512), "cpu", bf16_shadow=True)
    o = S.offsets
    # the three per-step weights (and their biases) must be contiguous: one GEMM, one weight-gradient GEMM
    a = o["attention.decoder_att.weight"]
    b = o["f_beta.weight"]
    c = o["decode_step.weight_hh"]
    assert a[0] + a[1] == b[0] and b[0] + b[1] == c[0]
    a, b, c = o["attention.decoder_att.bias"], o["f_beta.bias"], o["decode_step.bias_hh"]
    assert a[0] + a[1] == b[0] and b[0] + b[1] == c[0]
    assert o["init_h.weight"][0] + o["init_h.weight"][1] == o["init_c.weight"][0]
    assert all(v[0] % 8 == 0 for v in o.values())
    assert sum(v[1] for v in o.values()) == 4976117          # decoder parameter count at V=500
    S.f32("fc.bias").fill_(1.5)
    S.sync_shadow()
    assert S.w("fc.bias").dtype == torch.bfloat16 and float(S.w("fc.bias")[3]) == 1.5


def test_timing_signal_table_matches_oracle():
    from latex_ocr_b200.encoder import timing_signal_table
    from oracle import ref_model as rm
    t = timing_signal_table(512, 6, 30, "cpu")
    assert torch.equal(t, rm.timing_signal_nd(512, 6, 30).permute(1, 2, 0).contiguous())


def test_tf_decoder_variable_layout_and_adjacency():
    """TF-flavour Decoder (host side only): variables carry the TF names and shapes, load_tf_variables round-trips the oracle's
    dict, and the flat store keeps the blocks adjacent that the kernels address as one matrix (lo_tfdec_args comments)."""
    from latex_ocr_b200.tf_decoder import Decoder
    from oracle import ref_tf_model as tfm
    from util import Cfg
    V = 37
    cfg = Cfg(attn_cell_config={"num_units": 512, "dim_e": 256, "dim_o": 512, "dim_embeddings": 80}, decoding="beam_search", beam_size=3)
    dec = Decoder(cfg, V, V - 1, device="cpu", precision="bf16")
    p = tfm.init_params_tf(V, seed=4)
    sd = dec.state_dict()
    assert set(sd) == set(p)
    for k, v in p.items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
    dec.load_tf_variables(p)
    for k, v in p.items():
        assert torch.equal(dec.state_dict()[k], v), k
    S = dec.store
    off = {k: S.offsets[k][0] for k in S.offsets}
    n = {k: S.offsets[k][1] for k in S.offsets}
    for a, b in (("att_h.kernel", "o_W_h"), ("W_c_0", "W_h_0"), ("W_h_0", "W_o_0"), ("b_c_0", "b_h_0"), ("b_h_0", "b_o_0"),
                 ("embedding_table", "start_token")):
        assert off[a] + n[a] == off[b], (a, b)
    # storage is [out][in]: the TF-shaped parameter is a transposed view of it
    assert torch.equal(S.f32("lstm.kernel").t(), p["lstm.kernel"])
    assert dec._tiles == 3 and dec.max_length_formula == 150
    with pytest.raises(_lib.LatexOcrB200Error):
        dec.decode(torch.zeros(1, 4, 512))                      # CPU tensor: refused, no fallback


def test_encoder_variants_state_dict_surface():
    """'vanilla' and 'cnn' stacks (seq2seq_torch.py:31-86) expose the reference's state_dict keys and OIHW shapes."""
    from latex_ocr_b200.encoder import EncoderCNN
    from oracle import ref_model as rm
    from util import Cfg
    for variant in ("vanilla", "cnn"):
        pe, _ = rm.init_params(10, seed=1, encoder_cnn=variant)
        enc = EncoderCNN(Cfg(encoder_cnn=variant), device="cpu", precision="fp32")
        sd = enc.state_dict()
        assert set(sd) == set(pe)
        for k, v in pe.items():
            assert tuple(sd[k].shape) == tuple(v.shape), (variant, k)
        enc.load_state_dict(pe)
        assert all(torch.equal(enc.state_dict()[k], v) for k, v in pe.items())
    assert EncoderCNN(Cfg(encoder_cnn="cnn"), device="cpu", precision="fp32").out_hw(128, 512) == (15, 62)
    with pytest.raises(NotImplementedError):
        EncoderCNN(Cfg(encoder_cnn="resnet"), device="cpu")


def test_philox_host_mirror_known_answers():
    """Random123 kat_vectors for philox4x32-10 (the device generator in csrc/lo_common.cuh is the same ten rounds)."""
    from latex_ocr_b200 import philox
    kat = [((0, 0, 0, 0, 0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 6, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for args, want in kat:
        assert tuple(int(x) for x in philox.philox4x32_10(*args)) == want
    m = philox.dropout_multipliers(99, 0, 4, 5, 512, 0.5)
    assert m.shape == (4, 5, 512) and set(m.reshape(-1).tolist()) == {0.0, 2.0} and abs(m.mean() - 1.0) < 0.05
