"""cuBLAS DGEMM 8192^3 under ncu: what DMMA pipe utilisation does the library kernel reach?"""
import torch
a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda"); b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
for _ in range(3): torch.matmul(a, b)
torch.cuda.synchronize()
