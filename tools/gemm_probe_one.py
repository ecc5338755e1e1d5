import importlib, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
_lib = importlib.import_module("deep-spectral-segmentation_b200._lib")
lib = _lib.load()
dev = torch.device("cuda:0")
M = 32 * 901
N, K, bn, st = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
A = torch.randn(M, K, device=dev).half(); W = (torch.randn(N, K, device=dev) * 0.05).half()
bias = torch.zeros(N, device=dev); out = torch.empty(M, N, device=dev, dtype=torch.float16)
for _ in range(4):
    _lib.check(lib.dss_debug_gemm_cfg(A.data_ptr(), W.data_ptr(), bias.data_ptr(), out.data_ptr(), M, N, K, bn, st, _lib.stream_ptr()))
torch.cuda.synchronize()
