"""TensorFlow-flavour decoder — drop-in for the path of model/decoder.py:15-72 (``Decoder(config, n_tok, id_end)``),
model/components/attention_mechanism.py, attention_cell.py, greedy_decoder_cell.py and beam_search_decoder_cell.py on
the sm_100a kernels (csrc/lo_tfdecoder.cuh) through the C ABI (``lo_tfdec_*``).

The reference builds a TF graph; here the same three things the graph offers are methods:
  * ``train_logits(enc, formula)``            -> pred_train  [N,T,V]                      (decoder.py:48-57)
  * ``loss_and_backward(enc, formula, len)``  -> masked CE of img2seq.py:68-71 + every gradient (+ d enc)
  * ``decode(enc)``                           -> DecoderOutput(logits=None, ids)          (decoder.py:59-70; greedy ids [N,time],
                                                 beam ids [N,time,beam] like the TF cell before img2seq.py:241 transposes them)
``__call__(img, formula, dropout)`` returns ``(train_logits, DecoderOutput)`` like the reference.

Variables keep the TF names and shapes (``embedding_table`` [V,E], ``start_token`` [E], ``att_img.kernel`` [C,A],
``att_h.kernel`` [D,A] (tf.layers.dense inside compute_attention), ``att_beta`` [A], ``lstm.kernel`` [E+O+D,4D], ``lstm.bias``,
``o_W_c``, ``o_W_h``, ``y_W_o``, ``W_{c,h,o}_0``, ``b_{c,h,o}_0``) as views of a flat store whose memory is [out][in]
(K-major for the forward GEMMs).
"""
import collections
import ctypes
import math

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, stream_ptr
from .params import FlatStore, LRUCache, ParamHolder

DecoderOutput = collections.namedtuple("DecoderOutput", ("logits", "ids"))      # greedy_decoder_cell.py:5-6


def _make_args_struct():
    with open(_lib.HEADER) as f:
        text = f.read()

    class TfDecArgs(ctypes.Structure):
        _fields_ = _lib._parse_struct(text, "lo_tfdec_args")

    return TfDecArgs


TfDecArgs = _make_args_struct()
_bound = False


def _bind():
    global _bound
    L = _lib.lib()
    if _bound:
        return L
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    P = ctypes.POINTER(TfDecArgs)
    L.lo_sizeof_tfdec_args.restype = i64
    if L.lo_sizeof_tfdec_args() != ctypes.sizeof(TfDecArgs):
        raise _lib.LatexOcrB200Error("lo_tfdec_args layout mismatch: library %d bytes, ctypes mirror %d bytes — rebuild"
                                     % (L.lo_sizeof_tfdec_args(), ctypes.sizeof(TfDecArgs)))
    L.lo_tfdec_workspace_bytes.restype = i64
    L.lo_tfdec_workspace_bytes.argtypes = [P]
    for name, argtypes in (("lo_tfdec_forward", [P, i32, vp]), ("lo_tfdec_backward", [P, vp]),
                           ("lo_tfdec_greedy", [P, i64, i32, vp, vp, vp]), ("lo_tfdec_beam", [P, i64, i32, vp, vp, vp, vp, vp]),
                           ("lo_tfdec_beam_div", [P, i64, i32, vp, vp, vp, vp, ctypes.c_float, ctypes.c_float, vp, vp, vp])):
        fn = getattr(L, name)
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int
    _bound = True
    return L


def tf_decoder_specs(V, C, A, D, O, E):
    """Flat order: att_h.kernel^T and o_W_h^T are adjacent (one GEMM projects h_t onto both), the three initial-state
    matrices and their biases likewise, and start_token follows embedding_table (row V of the token table)."""
    return [
        ("att_img.kernel", (A, C)),
        ("att_h.kernel", (A, D)), ("o_W_h", (O, D)),
        ("att_beta", (A,)),
        ("lstm.kernel", (4 * D, E + O + D)), ("lstm.bias", (4 * D,)),
        ("o_W_c", (O, C)), ("y_W_o", (V, O)),
        ("W_c_0", (D, C)), ("W_h_0", (D, C)), ("W_o_0", (O, C)), ("b_c_0", (D,)), ("b_h_0", (D,)), ("b_o_0", (O,)),
        ("embedding_table", (V, E)), ("start_token", (E,)),
    ]


class Decoder(nn.Module):
    def __init__(self, config, n_tok, id_end, device="cuda", precision=None, impl=None, channels=512):
        super().__init__()
        self._config = config
        cell = getattr(config, "attn_cell_config", None) or {}
        self.D = int(cell.get("num_units", 512))
        self.A = int(cell.get("dim_e", 256))
        self.O = int(cell.get("dim_o", 512))
        self.E = int(cell.get("dim_embeddings", 80))
        self.C = channels
        self.V = int(n_tok)
        self._n_tok, self._id_end = int(n_tok), int(id_end)
        self.decoding = getattr(config, "decoding", "greedy")
        if self.decoding not in ("greedy", "beam_search"):
            raise NotImplementedError("decoding=%r" % (self.decoding,))
        self._tiles = 1 if self.decoding == "greedy" else int(getattr(config, "beam_size", 2))        # decoder.py:22
        # diversity penalty of beam_search_decoder_cell.py:258-287 (off in the shipped config: model.json:15-16 -> gamma 1, prob 0)
        self.div_gamma = float(getattr(config, "div_gamma", 1) or 1)
        self.div_prob = float(getattr(config, "div_prob", 0) or 0)
        self._div_state = None
        self.max_length_formula = int(getattr(config, "max_length_formula", 150))
        self.precision = precision or getattr(config, "precision", "bf16")
        self.impl = impl if impl is not None else getattr(config, "conv_impl", "tc" if self.precision == "bf16" else "simt")
        self.tdtype = torch.float32 if self.precision == "fp32" else torch.bfloat16
        self.store = FlatStore(tf_decoder_specs(self.V, self.C, self.A, self.D, self.O, self.E), device,
                               bf16_shadow=(self.precision == "bf16"))
        S = self.store
        bind = ParamHolder.bind
        for holder, sname in (("att_img", "att_img.kernel"), ("att_h", "att_h.kernel")):
            h = ParamHolder()
            h.bind("kernel", S, sname, permute=(1, 0))
            setattr(self, holder, h)
        self.lstm = ParamHolder()
        self.lstm.bind("kernel", S, "lstm.kernel", permute=(1, 0))
        self.lstm.bind("bias", S, "lstm.bias")
        for n in ("o_W_c", "o_W_h", "y_W_o", "W_c_0", "W_h_0", "W_o_0"):
            bind(self, n, S, n, permute=(1, 0))
        for n in ("att_beta", "b_c_0", "b_h_0", "b_o_0", "embedding_table", "start_token"):
            bind(self, n, S, n)
        self.reset_parameters()
        self._ws = LRUCache()     # bounded: see params.LRUCache
        self._shadow_fresh = False

    # tf.get_variable / tf.layers.dense default: glorot_uniform; LSTMCell bias zeros; embeddings decoder.py:98-105
    def reset_parameters(self):
        def glorot_(p, fan_in, fan_out):
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            p.uniform_(-lim, lim)

        with torch.no_grad():
            for p in (self.att_img.kernel, self.att_h.kernel, self.lstm.kernel, self.o_W_c, self.o_W_h, self.y_W_o, self.W_c_0,
                      self.W_h_0, self.W_o_0):
                glorot_(p, p.shape[0], p.shape[1])
            glorot_(self.att_beta, self.A, 1)
            for p in (self.b_c_0, self.b_h_0, self.b_o_0):
                glorot_(p, p.shape[0], p.shape[0])
            self.lstm.bias.zero_()
            for p in (self.embedding_table, self.start_token):
                p.uniform_(-1.0, 1.0)
                p.copy_(torch.nn.functional.normalize(p, dim=-1))
        self._shadow_fresh = False

    def load_tf_variables(self, d):
        """d: name -> tensor in the TF shapes (module docstring); ``att_beta`` may be [A,1]."""
        sd = self.state_dict()
        with torch.no_grad():
            for k, v in d.items():
                sd[k].copy_(v.reshape(sd[k].shape))
        self._shadow_fresh = False

    def _load_from_state_dict(self, *a, **k):
        super()._load_from_state_dict(*a, **k)
        self._shadow_fresh = False

    def sync_shadow(self):
        if not self._shadow_fresh:
            self.store.sync_shadow()
            self._shadow_fresh = True

    # ---------------------------------------------------------------------------------------------
    def _args(self, enc, B, T, R, rows_per_img=1):
        L = _bind()
        key = (B, T, R, rows_per_img)
        ws = self._ws.get(key)
        dev = self.store.device
        if ws is None:
            ws = {"ldl": (self.V + 63) // 64 * 64}
            ws["logits"] = torch.zeros(T, B, ws["ldl"], dtype=torch.float32, device=dev)
            ws["alphas"] = torch.zeros(B, T, R, dtype=torch.float32, device=dev)
            ws["loss"] = torch.zeros(4, dtype=torch.float32, device=dev)
            ws["denc"] = torch.zeros(B // rows_per_img, R, self.C, dtype=torch.float32, device=dev)
            ws["formula"] = torch.zeros(B, T, dtype=torch.int64, device=dev)
            ws["lengths"] = torch.zeros(B, dtype=torch.int32, device=dev)
            self._ws[key] = ws
        S = self.store
        a = TfDecArgs()
        a.B, a.T, a.R, a.C, a.A, a.D, a.O, a.E, a.V = B, T, R, self.C, self.A, self.D, self.O, self.E, self.V
        a.dt = _lib.LO_F32 if self.precision == "fp32" else _lib.LO_BF16
        a.impl = _lib.LO_IMPL_TC if (self.impl == "tc" and self.precision == "bf16") else _lib.LO_IMPL_SIMT
        a.ldl = ws["ldl"]
        a.rows_per_img = rows_per_img
        a.enc = enc.data_ptr()
        a.formula, a.formula_stride = ws["formula"].data_ptr(), ws["formula"].stride(0)
        a.lengths = ws["lengths"].data_ptr()
        W, F, G = (lambda n: S.w(n).data_ptr()), (lambda n: S.f32(n).data_ptr()), (lambda n: S.g(n).data_ptr())
        a.w_img, a.w_cat2, a.beta = W("att_img.kernel"), W("att_h.kernel"), F("att_beta")
        a.w_lstm, a.b_lstm, a.w_oc, a.w_y = W("lstm.kernel"), F("lstm.bias"), W("o_W_c"), W("y_W_o")
        a.w_init, a.b_init, a.emb = W("W_c_0"), F("b_c_0"), W("embedding_table")
        a.g_w_img, a.g_w_cat2, a.g_beta = G("att_img.kernel"), G("att_h.kernel"), G("att_beta")
        a.g_w_lstm, a.g_b_lstm, a.g_w_oc, a.g_w_y = G("lstm.kernel"), G("lstm.bias"), G("o_W_c"), G("y_W_o")
        a.g_w_init, a.g_b_init, a.g_emb = G("W_c_0"), G("b_c_0"), G("embedding_table")
        for k in ("logits", "alphas", "loss", "denc"):
            setattr(a, k, ws[k].data_ptr())
        if "ws" not in ws:
            ws["ws"] = torch.zeros(int(L.lo_tfdec_workspace_bytes(ctypes.byref(a))), dtype=torch.uint8, device=dev)
        a.ws = ws["ws"].data_ptr()
        ws["args"] = a
        return L, ws, a

    def _enc(self, enc):
        if not enc.is_cuda:
            raise _lib.LatexOcrB200Error("Decoder runs on CUDA tensors only (no CPU fallback)")
        if enc.dim() == 4:                                   # attention_mechanism.py:21-26: [N,H,W,C] -> [N,H*W,C]
            enc = enc.reshape(enc.shape[0], enc.shape[1] * enc.shape[2], enc.shape[3])
        return enc.to(self.tdtype).contiguous()

    def _keep(self, a, ws, keep_h, keep_o, T, B):
        """Optional dropout multipliers [N,T,D] / [N,T,O] (tf.nn.dropout semantics: 0 or 1/keep_prob) -> time-major."""
        for name, k, width in (("keep_h", keep_h, self.D), ("keep_o", keep_o, self.O)):
            if k is None:
                continue
            ws[name] = k.to(self.store.device, torch.float32).permute(1, 0, 2).contiguous()
            assert ws[name].shape == (T, B, width)
            setattr(a, name, ws[name].data_ptr())

    def run_forward(self, enc, formula, lengths=None, keep_h=None, keep_o=None, n_words=None):
        """enc [N,R,C] or [N,H,W,C]; formula int64 [N,T]; lengths [N] (incl. END) enables the loss.  Returns the workspace dict
        (``logits`` [T,N,ldl] time-major, ``alphas`` [N,T,R], ``loss``)."""
        enc = self._enc(enc)
        B, R, _ = enc.shape
        T = formula.shape[1]
        self.sync_shadow()
        L, ws, a = self._args(enc, B, T, R)
        ws["enc"] = enc
        ws["formula"].copy_(formula)
        with_loss = lengths is not None
        if with_loss:
            lens = torch.as_tensor(lengths, dtype=torch.int32)
            ws["lengths"].copy_(lens)
            # data parallel: n_words = token count of the GLOBAL batch, so that summing rank gradients gives the global mean
            a.inv_n_words = 1.0 / float(n_words if n_words is not None else int(lens.sum()))
        self._keep(a, ws, keep_h, keep_o, T, B)
        check(L.lo_tfdec_forward(ctypes.byref(a), 1 if with_loss else 0, stream_ptr()))
        return ws

    def run_backward(self, ws):
        check(_bind().lo_tfdec_backward(ctypes.byref(ws["args"]), stream_ptr()))
        return ws["denc"]

    def train_logits(self, enc, formula, keep_h=None, keep_o=None):
        with torch.no_grad():
            ws = self.run_forward(enc, formula, None, keep_h, keep_o)
            return ws["logits"][:, :, :self.V].permute(1, 0, 2).contiguous()

    def loss_and_backward(self, enc, formula, lengths, keep_h=None, keep_o=None, n_words=None):
        """Returns (loss tensor [4] on the device: mean CE, mean CE, 0, n_words; d loss / d enc [N,R,C] fp32).  Parameter
        gradients land in ``self.store.grad`` (the ``.grad`` of every parameter)."""
        with torch.no_grad():
            ws = self.run_forward(enc, formula, lengths, keep_h, keep_o, n_words)
            denc = self.run_backward(ws)
            return ws["loss"], denc

    def decode(self, enc, max_steps=None, div_u=None, return_attention=False):
        """dynamic_decode(decoder_cell, max_length_formula + 1) (decoder.py:70): at most max_length_formula + 2 steps, stops when
        every row / beam has emitted END.  Returns DecoderOutput(logits=None, ids).  ``div_u`` (tests): injected uniforms
        [steps, N*beam, V] for the diversity penalty; ``return_attention`` (greedy): also the attention weights [N, steps, R]
        (the tensor attention_mechanism.py:96-121 hands to visualize_attention.py)."""
        enc = self._enc(enc)
        N, R, _ = enc.shape
        beam = self._tiles
        steps = int(max_steps) if max_steps is not None else self.max_length_formula + 2
        B = N * beam
        self.sync_shadow()
        with torch.no_grad():
            L, ws, a = self._args(enc, B, steps, R, rows_per_img=beam)
            ws["enc"] = enc
            dev = self.store.device
            if self.decoding == "greedy":
                tokens = torch.zeros(N, steps, dtype=torch.int64, device=dev)
                fin = torch.zeros(N, steps, dtype=torch.int32, device=dev)
                check(L.lo_tfdec_greedy(ctypes.byref(a), self._id_end, steps, tokens.data_ptr(), fin.data_ptr(), stream_ptr()))
                done = fin.bool().all(dim=0)                            # dynamic_decode.py:38-40: loop ends once all rows finished
                n = int(torch.nonzero(done)[0]) + 1 if bool(done.any()) else steps
                if return_attention:
                    return DecoderOutput(None, tokens[:, :n]), ws["alphas"][:, :n].float().clone()
                return DecoderOutput(None, tokens[:, :n])
            ids = torch.zeros(N, steps, beam, dtype=torch.int64, device=dev)
            parents = torch.zeros_like(ids)
            fin = torch.zeros(N, steps, beam, dtype=torch.int32, device=dev)
            logp = torch.zeros(N, beam, dtype=torch.float32, device=dev)
            u_dev = None
            div_on = not (self.div_gamma == 1 or self.div_prob == 0)
            if div_on and div_u is not None:
                u_dev = torch.as_tensor(div_u, dtype=torch.float32).to(dev).contiguous()
                if tuple(u_dev.shape) != (steps, B, self.V):
                    raise ValueError("div_u must be [steps, N * beam, V]")
            elif div_on:
                if self._div_state is None:
                    self._div_state = torch.tensor([torch.initial_seed() & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64, device=dev)
                self._div_state[1] += 1                                  # a new Philox stream per decode call
            check(L.lo_tfdec_beam_div(ctypes.byref(a), self._id_end, steps, ids.data_ptr(), parents.data_ptr(), fin.data_ptr(),
                                      logp.data_ptr(), self.div_gamma, self.div_prob, u_dev.data_ptr() if u_dev is not None else None,
                                      self._div_state.data_ptr() if (div_on and u_dev is None) else None, stream_ptr()))
            done = fin.bool().all(dim=2).all(dim=0)
            n = int(torch.nonzero(done)[0]) + 1 if bool(done.any()) else steps
            # finalize (beam_search_decoder_cell.py:189-250) gathers with the unchanged initial parents -> identity
            return DecoderOutput(None, ids[:, :n])

    def forward(self, img, formula=None, dropout=1.0):
        """Decoder.__call__(img, formula, dropout) (decoder.py:24): (pred_train, pred_test).  dropout is the TF keep_prob; the
        shipped training config uses 1 (training.json:7) — other values draw masks here."""
        keep_h = keep_o = None
        train = None
        if formula is not None:
            if dropout is not None and float(dropout) < 1.0:
                N, T = formula.shape
                kp = float(dropout)
                dev = self.store.device
                keep_h = (torch.rand(N, T, self.D, device=dev) < kp).float() / kp
                keep_o = (torch.rand(N, T, self.O, device=dev) < kp).float() / kp
            train = self.train_logits(img, formula, keep_h, keep_o)
        return train, self.decode(img)
