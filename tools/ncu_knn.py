"""One launch of the kNN kernel at the Pool s3 site shape inside a cudaProfilerStart/Stop range (for `ncu --set full
--profile-from-start off`; diagnostic)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from nextou_b200 import ops

B, N, M, C, k = 1, 10752, 1344, 264, 28
x = torch.randn(B * N, C, device="cuda")
y = torch.randn(B * M, C, device="cuda")
rp = torch.randn(1, N, M, device="cuda") * 0.1
xn, sqx = ops.knn_normalize(x, B, N)
yn, sqy = ops.knn_normalize(y, B, M)
ops.knn_topk(xn, sqx, yn, sqy, rp, k, 1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ops.knn_topk(xn, sqx, yn, sqy, rp, k, 1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
