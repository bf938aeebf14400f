"""Why does the cfg #5 probe inside bench.py report a lower beam-5 rate than tools/decode_bench.py?  Same process set-up as bench.py
(train-built model, a few graph-captured train steps), then the probe twice, then per-batch timings of the beam mode."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench_support as bs
from latex_ocr_b200.img2seq import Img2SeqModel
from latex_ocr_b200.data import SimpleVocab
from latex_ocr_b200 import data as lod

class Cfg:
    encoder_cnn = "vanilla"; positional_embeddings = True; lr_init = 1e-3; lr_method = "adam"; cuda_graph = True
m = Img2SeqModel(Cfg(), vocab=SimpleVocab(500), device="cuda:0", precision="bf16", impl="tc")
m.build_train(); m.train_mode(True)
img, formula = bs.synthetic_batch(64, 128, 512, 500, 150, seed=1234)
img, formula = img.cuda(), formula.cuda()
for _ in range(4):
    m.train_step(img, formula)
torch.cuda.synchronize()
for k in range(2):
    r = bs.decode_probe(m, V=500, oracle_sample=False)
    print("probe call %d: greedy %.0f tok/s (%.3f s)  beam5 %.0f tok/s (%.3f s)" % (k, r["greedy_tok_s"], r["greedy_seconds"], r["beam5_tok_s"], r["beam5_seconds"]), flush=True)
print("decoder ws entries:", len(m.decoder._ws), "maxsize", m.decoder._ws.maxsize, "| encoder ws entries:", len(m.encoder._ws), "maxsize", m.encoder._ws.maxsize)
print("mem allocated %.1f GB reserved %.1f GB" % (torch.cuda.memory_allocated() / 1e9, torch.cuda.memory_reserved() / 1e9))
# per-batch timing of the beam mode
rng = np.random.RandomState(5)
m._config.decoding = "beam_search"; m._config.beam_size = 5; m._config.max_length_formula = 150
for W in (64, 256, 1024):
    b = [np.full((64, W, 1), 255, np.uint8) for _ in range(51)]
    x = torch.from_numpy(lod.pad_batch_images(b)).permute(0, 3, 1, 2).contiguous()
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ids, _ = m._decode_ids(x)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        from latex_ocr_b200 import decode
        t2 = time.perf_counter()
        ids2, _ = decode.beam_decode(m, x, 498, 499, 5, 150)
        torch.cuda.synchronize(); t3 = time.perf_counter()
        print("W %4d rep %d: _decode_ids %.1f ms | decode.beam_decode %.1f ms" % (W, rep, (t1 - t0) * 1e3, (t3 - t2) * 1e3), flush=True)
