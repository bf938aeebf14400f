import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from latex_ocr_b200 import _lib
L = _lib.lib()
M, N, K = 64, 3072, 512
A = torch.randn(M, K, device="cuda").bfloat16(); W = torch.randn(N, K, device="cuda").bfloat16()
C = torch.zeros(M, N, device="cuda"); b = torch.zeros(N, device="cuda")
for _ in range(20):
    _lib.check(L.lo_gemm(_lib.ptr(A), 1, _lib.ptr(W), 1, _lib.ptr(C), 0, M, N, K, K, 1, 1, K, N, 1, 0, 0, 0, _lib.ptr(b), 0, 0, 1, _lib.stream_ptr()))
torch.cuda.synchronize()
