"""torchrun --nproc-per-node N tests/manual/dp_check.py : N-rank data-parallel step == single-process step on the global batch.
Each rank takes its shard of a global batch; gradients are all-reduced (NCCL) bucket by bucket; rank 0 also runs the
global batch alone (no dist) and compares the post-Adam parameters and the pre-Adam (averaged) gradients."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import build_model  # noqa: E402
from oracle import ref_model as rm  # noqa: E402  (data + init only)
from latex_ocr_b200 import dist as lod  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
V, per = 60, 2
pe, pd = rm.init_params(V, seed=31)
img, formula = rm.synthetic_batch(per * world, 40, 72, V, 6, 6, seed=32)      # equal lengths: same padded T on every rank
T = formula.shape[1] - 1
for precision in ("fp32", "bf16"):
    m = build_model(V, pe, pd, precision)
    lod.attach(m)
    sh = slice(rank * per, (rank + 1) * per)
    loss = m._step_body(img[sh].cuda(), formula[sh].cuda(), [T] * per, None)
    torch.cuda.synchronize()
    gd = (m.decoder.store.grad / world).clone()
    ge = (m.encoder.store.grad / world).clone()
    losses = [torch.zeros(4, device="cuda") for _ in range(world)]
    dist.all_gather(losses, loss)
    if rank == 0:
        ref = build_model(V, pe, pd, precision)
        lref = ref._step_body(img.cuda(), formula.cuda(), [T] * (per * world), None)
        torch.cuda.synchronize()
        tol = 1e-4 if precision == "fp32" else 3e-2
        ed = (gd - ref.decoder.store.grad).abs().max().item() / ref.decoder.store.grad.abs().max().item()
        ee = (ge - ref.encoder.store.grad).abs().max().item() / ref.encoder.store.grad.abs().max().item()
        mean_loss = sum(l[0].item() for l in losses) / world
        print("[dp_check %s] world=%d  loss mean-of-ranks %.6f vs global %.6f | grad rel.err dec %.2e enc %.2e"
              % (precision, world, mean_loss, lref[0].item(), ed, ee), flush=True)
        assert abs(mean_loss - lref[0].item()) / abs(lref[0].item()) < tol
        assert ed < 10 * tol and ee < 10 * tol
dist.barrier()
if rank == 0:
    print("dp_check OK", flush=True)
dist.destroy_process_group()
