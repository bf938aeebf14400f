"""TensorFlow-flavour ``Img2SeqModel`` — the surface of model/img2seq.py on the sm_100a kernels: ``build_train`` /
``build_pred``, ``_run_train(config, train_set, val_set, epoch, lr_schedule)`` (:144-196), ``write_prediction`` /
``_run_evaluate`` semantics (:198-254, perplexity negated like :252), ``predict_batch(images)`` / ``predict(img)`` (:256-285).

Encoder: the same 6-conv stack as the torch flavour with the TF input normalisation (img - 128) / 128 fused into conv1
(model/encoder.py:26-27; SAME/VALID paddings and SAME pools coincide with the torch stack for even sizes).
Decoder: ``tf_decoder.Decoder`` (Genthial cell).  Loss: masked CE (img2seq.py:68-71).  Optimiser: Adam on the two flat stores
(tf.train.AdamOptimizer, :104; epsilon 1e-8 like the torch flavour), optional clip_by_global_norm (:116-121).
"""
import os
import time

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr
from .data import minibatches, pad_batch_formulas, pad_batch_images
from .encoder import EncoderCNN
from .tf_decoder import Decoder


class _Cfg:
    """config view that adds input_norm='tf' without mutating the caller's object"""

    def __init__(self, base):
        self._base = base
        self.input_norm = "tf"

    def __getattr__(self, k):
        return getattr(self._base, k)


class Img2SeqModel:
    def __init__(self, config, dir_output=None, vocab=None, device=None, precision=None, impl=None):
        self._config = config
        self._dir_output = dir_output
        self._vocab = vocab
        self.device = torch.device(device or getattr(config, "device", "cuda"))
        if self.device.type != "cuda":
            raise _lib.LatexOcrB200Error("latex_ocr_b200 runs on CUDA devices only (no CPU fallback)")
        self.precision = precision or getattr(config, "precision", "bf16")
        self.impl = impl if impl is not None else getattr(config, "conv_impl", "tc" if self.precision == "bf16" else "simt")
        self.encoder = self.decoder = None
        self.lr = None
        self.dist = None          # set by latex_ocr_b200.dist.attach(model): data-parallel gradient all-reduce
        self.last_epoch_stats = {}

    # img2seq.py:33-53 / :55-66: the TF graph is replaced by the two modules
    def build_train(self, config=None):
        config = config or self._config
        self._build()
        method = str(getattr(config, "lr_method", "adam")).lower()
        # img2seq.py:98-111: adam | adagrad | sgd | rmsprop, TF 1.12 update rules (csrc/lo_optim.cu:tf_optim_kernel)
        kinds = {"adam": 1, "sgd": 2, "adagrad": 3, "rmsprop": 4}
        if method not in kinds:
            raise NotImplementedError("Unknown method {}".format(method))          # img2seq.py:111
        self.lr_method, self._opt_kind = method, kinds[method]
        self.clip = float(getattr(config, "clip", -1))
        self.set_lr(float(getattr(config, "lr_init", 1e-3)))
        return self

    def build_pred(self, config=None):
        self._build()
        return self

    def _build(self):
        if self.encoder is None:
            self.encoder = EncoderCNN(_Cfg(self._config), device=self.device, precision=self.precision, impl=self.impl)
            self.decoder = Decoder(self._config, self._vocab.n_tok, self._vocab.id_end, device=self.device, precision=self.precision,
                                   impl=self.impl)

    def set_lr(self, lr):
        self.lr = float(lr)
        kind = getattr(self, "_opt_kind", 1)
        for m in (self.encoder, self.decoder):
            S = m.store
            fresh = S.m is None
            S.ensure_adam(lr)                     # slot buffers m (s1) / v (s2) + the device {step, lr} pair
            if fresh and kind == 3:
                S.m.fill_(0.1)                    # tf.train.AdagradOptimizer: initial_accumulator_value = 0.1
            elif fresh and kind == 4:
                S.m.fill_(1.0)                    # tf.train.RMSPropOptimizer: the rms slot starts at ones
            S.set_lr(lr)

    # ---------------------------------------------------------------------------------------------
    def _to_device_images(self, images):
        """list of HWC uint8 arrays (or an [N,H,W,1] array) -> CUDA uint8 [N,1,H,W] (pad_batch_images: 255 padding)."""
        if isinstance(images, torch.Tensor):
            img = images
        else:
            img = torch.from_numpy(pad_batch_images(images))
        if img.dim() == 4 and img.shape[-1] == 1:
            img = img.permute(0, 3, 1, 2)
        return img.contiguous().to(self.device, non_blocking=True)

    def compute_gradients(self, images, formulas, dropout=1.0):
        """Forward + loss + backward (+ the data-parallel all-reduce): images as above, formulas = list of id lists (padded here like
        _get_feed_dict img2seq.py:132-135) or an (int array [N,T], lengths [N]) pair.  Gradients land in the two flat stores;
        returns the device loss vector [mean CE, mean CE, 0, n_words]."""
        if isinstance(formulas, tuple):
            formula, length = formulas
        else:
            formula, length = pad_batch_formulas(formulas, self._vocab.id_pad, self._vocab.id_end)
        formula_t = torch.as_tensor(np.asarray(formula), dtype=torch.int64).to(self.device, non_blocking=True)
        length_t = torch.as_tensor(np.asarray(length), dtype=torch.int32)
        img = self._to_device_images(images)
        N = img.shape[0]
        enc = self.encoder.forward_raw(img, need_grad=True)
        keep_h = keep_o = None
        if dropout is not None and float(dropout) < 1.0:
            kp, T = float(dropout), formula_t.shape[1]
            keep_h = (torch.rand(N, T, self.decoder.D, device=self.device) < kp).float() / kp
            keep_o = (torch.rand(N, T, self.decoder.O, device=self.device) < kp).float() / kp
        n_words = None
        if self.dist is not None and self.dist.world_size > 1:
            # the loss is a mean over valid tokens: all-reduce the token count (one scalar), then the SUM of the rank
            # gradients is the gradient of the global mean
            import torch.distributed as tdist
            nw = torch.tensor([float(length_t.sum())], device=self.device)
            tdist.all_reduce(nw, group=self.dist.group)
            n_words = float(nw.item())
        loss, denc = self.decoder.loss_and_backward(enc, formula_t, length_t, keep_h, keep_o, n_words)
        if self.dist is not None:
            self.dist.reduce_async(self.decoder.store.grad)             # flies while the encoder backward runs
        self.encoder.backward_raw(tuple(img.shape), denc.view(N, enc.shape[1], enc.shape[2], enc.shape[3]))
        if self.dist is not None:
            self.dist.reduce_async(self.encoder.store.grad)
            self.dist.wait()
        return loss

    def apply_gradients(self):
        """optimizer.apply_gradients (img2seq.py:113-123): optional clip_by_global_norm, then Adam on both stores."""
        L = _lib.lib()
        scale = 1.0
        if self.clip > 0:                                                   # tf.clip_by_global_norm (img2seq.py:116-121)
            gn = float(torch.sqrt(self.encoder.store.grad.pow(2).sum() + self.decoder.store.grad.pow(2).sum()))
            scale = min(1.0, self.clip / max(gn, 1e-30))
        import ctypes
        fn = L.lo_tf_optim_step
        fn.restype = ctypes.c_int
        fn.argtypes = [ctypes.c_int] + [ctypes.c_void_p] * 5 + [ctypes.c_int64, ctypes.c_void_p] + [ctypes.c_float] * 4 + [ctypes.c_void_p]
        kind = getattr(self, "_opt_kind", 1)
        # TF defaults: Adam beta1 0.9, beta2 0.999, epsilon 1e-8 (outside the bias correction); RMSProp decay 0.9, epsilon 1e-10
        b1, b2, eps = (0.9, 0.9, 1e-10) if kind == 4 else (0.9, 0.999, 1e-8)
        for m in (self.decoder, self.encoder):
            S = m.store
            check(fn(kind, ptr(S.master), ptr(S.grad), ptr(S.m), ptr(S.v), ptr(S.shadow), S.numel, ptr(S.adam_state), b1, b2, eps,
                     float(scale), stream_ptr()))
            m._shadow_fresh = True                                          # the fused update refreshed the bf16 shadow

    def train_step(self, images, formulas, dropout=1.0):
        """One update (img2seq.py:163-170).  Returns the device loss vector [mean CE, mean CE, 0, n_words]."""
        loss = self.compute_gradients(images, formulas, dropout)
        self.apply_gradients()
        return loss

    def _run_train(self, config, train_set, val_set, epoch, lr_schedule):
        """img2seq.py:144-196."""
        batch_size = config.batch_size
        nbatches = (len(train_set) + batch_size - 1) // batch_size
        t0, n_img, losses = time.time(), 0, []
        for i, (img, formula) in enumerate(minibatches(train_set, batch_size)):
            if lr_schedule is not None:
                self.set_lr(lr_schedule.lr)
            loss = self.train_step(img, formula, dropout=getattr(config, "dropout", 1.0))
            losses.append(loss[0:1].clone())
            n_img += len(img)
            if lr_schedule is not None:
                lr_schedule.update(batch_no=epoch * nbatches + i)
        torch.cuda.synchronize()
        mean_loss = float(torch.cat(losses).mean()) if losses else float("nan")
        self.last_epoch_stats = {"loss": mean_loss, "perplexity_train": float(np.exp(mean_loss)),
                                 "images_per_s": n_img / max(time.time() - t0, 1e-9)}
        if val_set is None:
            return -float(np.exp(mean_loss))
        scores = self.evaluate(config, val_set)
        self.last_epoch_stats.update(scores)
        if lr_schedule is not None:
            lr_schedule.update(score=scores["perplexity"])
        return scores["perplexity"]

    def evaluate(self, config, test_set):
        """_run_evaluate / write_prediction (img2seq.py:198-254): perplexity = -exp(sum CE / n_words) from the teacher-forced
        pass with dropout off, text metrics of hypothesis 0 of the decoded ids."""
        from . import decode, metrics
        refs, hyps = [], []
        ce_words, n_words = 0.0, 0.0
        for imgs, formulas in minibatches(test_set, config.batch_size):
            img = self._to_device_images(imgs)
            formula, length = pad_batch_formulas(formulas, self._vocab.id_pad, self._vocab.id_end)
            enc = self.encoder.forward_raw(img)
            ws = self.decoder.run_forward(enc, torch.as_tensor(formula.astype(np.int64)).to(self.device),
                                          torch.as_tensor(length.astype(np.int32)))
            l = ws["loss"].tolist()
            ce_words += l[0] * l[3]
            n_words += l[3]
            ids = self.decoder.decode(enc).ids
            hyp0 = ids if ids.dim() == 2 else ids[:, :, 0]
            hyps += decode.truncate_end(hyp0.tolist(), self._vocab.id_end)
            refs += [list(map(int, f)) for f in formulas]
        scores = metrics.score(refs, hyps)
        scores["perplexity"] = -float(np.exp(ce_words / max(n_words, 1.0)))
        return scores

    def predict_batch(self, images):
        """img2seq.py:256-275: list over hypotheses (1 for greedy, beam_size for beam search) of N token-id lists, truncated at END."""
        from . import decode
        enc = self.encoder.forward_raw(self._to_device_images(images))
        ids = self.decoder.decode(enc).ids
        if ids.dim() == 2:
            return [decode.truncate_end(ids.tolist(), self._vocab.id_end)]
        return [decode.truncate_end(ids[:, :, k].tolist(), self._vocab.id_end) for k in range(ids.shape[2])]

    def predict(self, img):
        return [h[0] for h in self.predict_batch([img])]

    # checkpoints
    def state_dict(self):
        return {"encoder": self.encoder.state_dict(), "decoder": self.decoder.state_dict()}

    def load_state_dict(self, sd):
        self.encoder.load_state_dict(sd["encoder"])
        self.decoder.load_state_dict(sd["decoder"])

    def save(self, path=None):
        path = path or os.path.join(self._dir_output or ".", "model_tf.pt")
        torch.save({k: {n: t.detach().cpu().contiguous() for n, t in v.items()} for k, v in self.state_dict().items()}, path)
        return path

    def restore(self, path):
        self.load_state_dict(torch.load(path, map_location=self.device))
