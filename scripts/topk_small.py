"""One user-block sweep per CTA (37,888 x 1,000,000, r=128, k=100): the ncu --set full target of the top-k kernels."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from teamoflow_b200.mf._engine import new_storage
from teamoflow_b200.mf.matrix_factorization import score_topk
n_u, n_i, r, k = 37888, 1_000_000, 128, 100
dev = torch.device("cuda")
g = torch.Generator(device=dev); g.manual_seed(20245)
U = new_storage(n_u, r); U[:, :r] = torch.randn(n_u, r, generator=g, device=dev) / math.sqrt(r)
V = new_storage(n_i, r); V[:, :r] = torch.randn(n_i, r, generator=g, device=dev) / math.sqrt(r)
for _ in range(4):
    score_topk(U, V, r, k, False, 0)
torch.cuda.synchronize()
