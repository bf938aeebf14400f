import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import ref_model as rm
import bench_support as bs
print("affinity", len(os.sched_getaffinity(0)), "cpu_count", os.cpu_count(), "cpu.max", open("/sys/fs/cgroup/cpu.max").read().strip() if os.path.exists("/sys/fs/cgroup/cpu.max") else None, "chosen", bs.cpu_threads(), flush=True)
pe, pd = rm.init_params(500, seed=0)
img, formula = rm.synthetic_batch(4, 128, 512, 500, 20, 20, seed=1)
for th in (8, 16, 32, 64):
    torch.set_num_threads(th)
    rm.train_step(pe, pd, img, formula, {})
    t0 = time.perf_counter(); rm.train_step(pe, pd, img, formula, {}); dt = time.perf_counter() - t0
    print("threads", th, "B=4 T=20 step %.2f s" % dt, flush=True)
