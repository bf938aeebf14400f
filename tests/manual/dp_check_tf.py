"""torchrun --nproc-per-node N tests/manual/dp_check_tf.py : TF-flavour data parallel == single process on the global batch.
The masked loss is a mean over valid tokens, so the ranks first all-reduce the token count (img2seq_tf.compute_gradients);
the SUM of the rank gradients must equal the global-batch gradient.  Ragged lengths on purpose."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import Cfg  # noqa: E402
from latex_ocr_b200 import dist as lod  # noqa: E402
from latex_ocr_b200.data import SimpleVocab  # noqa: E402
from latex_ocr_b200.img2seq_tf import Img2SeqModel  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
V, per = 30, 2
rng = np.random.RandomState(5)
imgs = [rng.randint(0, 256, (32, 64, 1)).astype(np.uint8) for _ in range(per * world)]
forms = [list(rng.randint(0, V - 3, 2 + (i * 3) % 5)) for i in range(per * world)]
cfg = Cfg(attn_cell_config={"num_units": 512, "dim_e": 256, "dim_o": 512, "dim_embeddings": 80}, max_length_formula=10, decoding="greedy",
          batch_size=per, lr_init=1e-3)
from latex_ocr_b200.data import pad_batch_formulas  # noqa: E402
formula_all, len_all = pad_batch_formulas(forms, V - 2, V - 1)            # one padded length for every rank
for precision in ("fp32", "bf16"):
    torch.manual_seed(7)
    m = Img2SeqModel(cfg, vocab=SimpleVocab(V), device="cuda:%d" % local, precision=precision).build_train(cfg)
    lod.attach(m)                                                           # broadcasts rank 0's parameters
    sh = slice(rank * per, (rank + 1) * per)
    m.compute_gradients(imgs[sh], (formula_all[sh], len_all[sh]))
    torch.cuda.synchronize()
    gd, ge = m.decoder.store.grad.clone(), m.encoder.store.grad.clone()
    if rank == 0:
        ref = Img2SeqModel(cfg, vocab=SimpleVocab(V), device="cuda:%d" % local, precision=precision).build_train(cfg)
        ref.load_state_dict(m.state_dict())
        ref.compute_gradients(imgs, (formula_all, len_all))
        torch.cuda.synchronize()
        ed = (gd - ref.decoder.store.grad).abs().max().item() / ref.decoder.store.grad.abs().max().item()
        ee = (ge - ref.encoder.store.grad).abs().max().item() / ref.encoder.store.grad.abs().max().item()
        print("[dp_check_tf %s] world=%d grad rel.err dec %.2e enc %.2e" % (precision, world, ed, ee), flush=True)
        tol = 1e-4 if precision == "fp32" else 3e-2
        assert ed < 10 * tol and ee < 10 * tol
dist.barrier()
if rank == 0:
    print("dp_check_tf OK", flush=True)
dist.destroy_process_group()
