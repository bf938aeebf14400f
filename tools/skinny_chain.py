"""Latency of a dependent chain of per-step GEMMs (skinny_mma_kernel, programmatic dependent launch, CUDA graph) as a function of the
rows actually read (rows >= M are zero-filled without touching memory) and of the number of column tiles: tells how much of a launch
is the post-wait activation load."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from latex_ocr_b200 import _lib
L = _lib.lib()
K = 512
n_chain = 60
for N in (3072, 2048, 1024, 512):
    for M in (64, 32, 16):
        A = torch.randn(64, K, device="cuda").bfloat16(); W = torch.randn(N, K, device="cuda").bfloat16()
        C = torch.zeros(64, N, device="cuda"); b = torch.zeros(N, device="cuda")

        def chain():
            for _ in range(n_chain):
                _lib.check(L.lo_gemm(_lib.ptr(A), 1, _lib.ptr(W), 1, _lib.ptr(C), 0, M, N, K, K, 1, 1, K, N, 1, 0, 0, 0, _lib.ptr(b), 0, 0, 1,
                                     _lib.stream_ptr()))
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            chain(); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                chain()
        torch.cuda.synchronize()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        print("N %4d (%3d CTAs)  M %2d rows read: %.2f us per launch" % (N, N // 16, M, e0.elapsed_time(e1) / 5 / n_chain * 1e3), flush=True)
