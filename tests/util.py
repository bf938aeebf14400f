"""Shared helpers for the parity tests (the oracle is imported ONLY from tests/)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def relerr(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


class Cfg:
    def __init__(self, **kw):
        self.encoder_cnn = "vanilla"
        self.positional_embeddings = True
        self.lr_init = 1e-3
        self.lr_method = "adam"
        self.batch_size = 4
        self.n_epochs = 1
        self.__dict__.update(kw)


def build_model(V, pe, pd, precision="fp32", positional=True, train=False, impl="simt", graph=False):
    from latex_ocr_b200.img2seq import Img2SeqModel
    m = Img2SeqModel(Cfg(positional_embeddings=positional, cuda_graph=graph), n_tok=V, device="cuda",
                     precision=precision, impl=impl)
    m.build_train()
    m.encoder.load_state_dict(pe)
    m.decoder.load_state_dict(pd)
    m.train_mode(train)
    return m


def grads_as_reference_layout(module):
    """{state_dict name: gradient tensor in the reference layout (OIHW for convs)}"""
    return {k: p.grad.detach().float().cpu() for k, p in module.named_parameters()}
