"""L2 persistence experiment: time 150 consecutive attention forward launches (one decode pass, same att1/enc every launch) with
(a) the default cache hints, (b) no hints, (c) an access-policy window pinning enc (57 MB), (d) pinning att1, (e) window over both."""
import ctypes, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_support as bs
from latex_ocr_b200 import _lib
from latex_ocr_b200.img2seq import Img2SeqModel
from latex_ocr_b200.data import SimpleVocab

B, T, R, A, C = 64, 150, 868, 512, 512
class Cfg:
    encoder_cnn = "vanilla"; positional_embeddings = True; lr_init = 1e-3; lr_method = "adam"; cuda_graph = False
m = Img2SeqModel(Cfg(), vocab=SimpleVocab(500), device="cuda:0", precision="bf16", impl="tc").build_train()
m.train_mode(True)
img, formula = bs.synthetic_batch(B, 128, 512, 500, T, seed=1234)
m.train_step(img.cuda(), formula.cuda())
torch.cuda.synchronize()
L = _lib.lib()
dec, enc = m.decoder, m.encoder
ws = dec._ws[[k for k in dec._ws if k[0] == B and k[1] == T][0]]
t, a = ws["t"], ws["args"]
enc_out = enc._ws[(B, 128, 512)]["out"].view(B, R, C)
O1 = A + C + 4 * 512
side = torch.cuda.Stream()                     # a real (non-default) stream: the access-policy window is a stream attribute
torch.cuda.set_stream(side)
st = _lib.stream_ptr()
dt = _lib.LO_BF16

def att_steps():
    for s in range(T):
        _lib.check(L.lo_attention_forward(_lib.ptr(t["att1"]), _lib.ptr(enc_out), dt, _lib.ptr(t["out1"][s]), O1, a.w_full,
                                          ctypes.c_void_p(t["alphas"].data_ptr() + s * R * 4), T * R, _lib.ptr(t["ctx"][s]), None, 0, None,
                                          B, R, A, C, _lib.ptr(t["work"]), st))

def run(tag):
    us = bs._time_ms(att_steps, 3) / T * 1e3
    print(json.dumps({"case": tag, "us_per_launch": round(us, 2), "GBps": round((B * R * (A + C) * 2 + B * R * 4) / us / 1e3, 1)}), flush=True)

run("default hints (enc evict_last, att1 evict_first)")
_lib.set_option("att_policy_enc", 3); _lib.set_option("att_policy_att1", 3)
run("no hints")
nbytes = B * R * C * 2
for tag, ptr, n, ratio in (("window enc 57MB", enc_out.data_ptr(), nbytes, 1.0), ("window att1 57MB", t["att1"].data_ptr(), nbytes, 1.0),
                           ("window enc 57MB ratio 0.6", enc_out.data_ptr(), nbytes, 0.6)):
    _lib.check(L.lo_set_l2_window(ctypes.c_void_p(ptr), n, ratio, st))
    print(L.lo_last_error().decode())
    run(tag)
    _lib.check(L.lo_set_l2_window(None, 0, 0.0, st))
_lib.set_option("att_policy_enc", 1); _lib.set_option("att_policy_att1", 3)
_lib.check(L.lo_set_l2_window(ctypes.c_void_p(t["att1"].data_ptr()), nbytes, 1.0, st))
run("enc evict_last hint + window att1")
_lib.check(L.lo_set_l2_window(None, 0, 0.0, st))
