"""Where the decoder's time goes: the decoder forward and backward at the bench workload, each captured in a CUDA graph, with parts of
the schedule switched off (lo_set_option("dbg_skip", mask): results are garbage, only the timing is meaningful)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench_support as bs
from latex_ocr_b200 import _lib
from latex_ocr_b200.img2seq import Img2SeqModel
from latex_ocr_b200.data import SimpleVocab

B, T = 64, 150
class Cfg:
    encoder_cnn = "vanilla"; positional_embeddings = True; lr_init = 1e-3; lr_method = "adam"; cuda_graph = False
m = Img2SeqModel(Cfg(), vocab=SimpleVocab(500), device="cuda:0", precision="bf16", impl="tc")
m.build_train(); m.train_mode(True)
img, formula = bs.synthetic_batch(B, 128, 512, 500, T, seed=1234)
img, formula = img.cuda(), formula.cuda()
for _ in range(2):
    m.train_step(img, formula)
torch.cuda.synchronize()
L = _lib.lib()
dec = m.decoder
key = [k for k in dec._ws if k[0] == B and k[1] == T][0]
a = dec._ws[key]["args"]


def timed(fn, name, iters=5):
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        fn(); fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            fn()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print("%-58s %8.3f ms" % (name, ms), flush=True)
    return ms


def fwd():
    _lib.check(L.lo_decoder_forward(ctypes.byref(a), 1, _lib.stream_ptr()))


def bwd():
    _lib.check(L.lo_decoder_backward(ctypes.byref(a), _lib.stream_ptr()))


res = {}
for mask, label in ((0, "all"), (8 | 1, "time loop only (no hoisted work)"), (8 | 1 | 2, "loop without the attention kernels"),
                    (8 | 1 | 4, "loop with ONLY the attention kernels")):
    _lib.set_option("dbg_skip", mask)
    res[("f", mask)] = timed(fwd, "forward  : " + label)
    res[("b", mask)] = timed(bwd, "backward : " + label)
_lib.set_option("dbg_skip", 0)
fwd(); bwd(); torch.cuda.synchronize()
print("per step (us): fwd loop %.1f, att-only %.1f, small-only %.1f | bwd loop %.1f, att-only %.1f, small-only %.1f" % (
    res[("f", 9)] / T * 1e3, res[("f", 13)] / T * 1e3, res[("f", 11)] / T * 1e3,
    res[("b", 9)] / T * 1e3, res[("b", 13)] / T * 1e3, res[("b", 11)] / T * 1e3))
