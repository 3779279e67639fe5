import torch
m, n, k = 16384, 20480, 4096
A = torch.randn(k, m, device="cuda", dtype=torch.float64)
B = torch.randn(k, n, device="cuda", dtype=torch.float64)
for _ in range(3):
    C = A.t() @ B
torch.cuda.synchronize()
