"""(build the timing variant first: LO_LIB_DIR=_C_timing LO_NVCC_EXTRA=-DLO_ATT_TIMING python -m latex_ocr_b200.build)
Per-CTA timeline of the grid-barrier fused forward step kernel (dec_fuse=1; timing build LO_LIB_DIR=_C_timing, -DLO_ATT_TIMING),
time loop only, attention switched off.  Stamps: 0 entry, 1 weights issued, 2 past griddepcontrol.wait, 3 A operand landed,
4 cell done (before the grid barrier), 5 past the barrier, 6 A of phase 2 landed, 7 end."""
import ctypes, os, sys
os.environ.setdefault("LO_LIB_DIR", "_C_timing")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench_support as bs
from latex_ocr_b200 import _lib
from latex_ocr_b200.img2seq import Img2SeqModel
from latex_ocr_b200.data import SimpleVocab

B, T = 64, 150
class Cfg:
    encoder_cnn = "vanilla"; positional_embeddings = True; lr_init = 1e-3; lr_method = "adam"; cuda_graph = False
_lib.set_option("dec_fuse", 1)
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
buf = torch.zeros(1024 * 16, dtype=torch.int64, device="cuda")
for mask, lab in ((9 | 2, "smalls only"),):
    _lib.set_option("dbg_skip", mask)
    for rep in range(2):
        buf.zero_()
        _lib.check(L.lo_debug_buffer(_lib.ptr(buf)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.lo_decoder_forward(ctypes.byref(a), 1, _lib.stream_ptr()))
        e1.record(); torch.cuda.synchronize()
    print("forward loop, %s: %.2f us per step (eager launches)" % (lab, e0.elapsed_time(e1) / T * 1e3))
    x = buf.cpu().numpy().reshape(-1, 16)[:192].astype(np.float64)
    t0 = x[:, 0][x[:, 0] > 0].min()
    for k, nm in ((0, "entry"), (1, "weights issued"), (2, "past griddepcontrol.wait"), (3, "A operand landed"), (4, "cell done"),
                  (5, "past grid barrier"), (6, "A of phase 2 landed"), (7, "end")):
        v = (x[:, k] - t0) / 1e3
        v = v[x[:, k] > 0]
        if len(v):
            print("  %-28s %7.2f / %7.2f / %7.2f   (%d CTAs)" % (nm, v.min(), np.median(v), v.max(), len(v)))
_lib.set_option("dbg_skip", 0)
_lib.check(L.lo_debug_buffer(None))
