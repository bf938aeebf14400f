"""CPU (-m "not gpu"): the data-parallel path at world_size 2 over gloo.  Checks (i) GradSync's bucketed all-reduce and
(ii) the claim the design rests on: with equal per-rank batches and the same padded length, the
all-reduced (summed) rank gradients times 1/world equal the gradient of the loss on the global batch — using the
CPU oracle for the per-rank gradients (the GPU kernels are checked against the same oracle in test_gpu_parity.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from latex_ocr_b200.dist import GradSync
    from oracle import ref_model as rm
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    V = 30
    pe, pd = rm.init_params(V, seed=1)
    img, formula = rm.synthetic_batch(4, 32, 64, V, 5, 5, seed=2)       # global batch 4, equal lengths
    shard = slice(rank * 2, rank * 2 + 2)
    _, ge, gd, _ = rm.train_step({k: v.clone() for k, v in pe.items()}, {k: v.clone() for k, v in pd.items()},
                                 img[shard], formula[shard], {})
    sync = GradSync()
    flat_d = torch.cat([gd[k].reshape(-1) for k in sorted(gd)])        # decoder bucket
    flat_e = torch.cat([ge[k].reshape(-1) for k in sorted(ge)])        # encoder bucket
    sync.reduce_async(flat_d)
    sync.reduce_async(flat_e)
    sync.wait()
    flat_d /= sync.world_size
    flat_e /= sync.world_size
    # per-layer encoder buckets (GradSync.layer_hook): every element of the flat gradient is all-reduced exactly once, in
    # the order the encoder backward completes its layers (last conv first), small layers coalesced
    class FakeEnc:
        layers = (("0",), ("3",), ("6",), ("8",))
        ranges = {"0": (0, 8), "3": (8, 40), "6": (48, 400), "8": (448, 1000)}

        class store:
            grad = torch.full((1448,), float(rank + 1))

        def grad_range(self, idx):
            return self.ranges[idx]
    fe = FakeEnc()
    calls = []
    orig = sync.reduce_async
    sync.reduce_async = lambda t: (calls.append(t.numel()), orig(t))[1]
    hook = sync.layer_hook(fe, min_bucket=300)
    for idx in ("8", "6", "3", "0"):
        hook(idx)
    sync.wait()
    sync.reduce_async = orig
    assert calls == [1000, 400, 48], calls
    assert torch.equal(fe.store.grad, torch.full((1448,), 3.0))          # 1 + 2 over the two ranks, each element once
    if rank == 0:
        _, ge_full, gd_full, _ = rm.train_step({k: v.clone() for k, v in pe.items()}, {k: v.clone() for k, v in pd.items()},
                                               img, formula, {})
        want_d = torch.cat([gd_full[k].reshape(-1) for k in sorted(gd_full)])
        want_e = torch.cat([ge_full[k].reshape(-1) for k in sorted(ge_full)])
        out.put(((flat_d - want_d).abs().max().item() / want_d.abs().max().item(),
                 (flat_e - want_e).abs().max().item() / want_e.abs().max().item(), sync.world_size))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gradient_mean_equals_global_batch_gradient():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ed, ee, ws = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert ws == 2
    # the regulariser mean((1 - sum_t alpha)^2) and the CE mean are both means over per-rank rows -> exact averaging
    assert ed < 1e-4 and ee < 1e-4, (ed, ee)
