"""One bounded slab pass for ncu launch lists: python scripts/probe_one.py G [bounded]"""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from teamoflow_b200.mf import dist as tdist
from teamoflow_b200.mf._engine import new_storage
from teamoflow_b200.mf.matrix_factorization import score_topk
G_ = int(sys.argv[1]); bounded = len(sys.argv) > 2 and sys.argv[2] == "bounded"
n_u, n_i, r, k = 1_000_000, 1_000_000, 128, 100
dev = torch.device("cuda")
g = torch.Generator(device=dev); g.manual_seed(20245)
U = new_storage(n_u, r); U[:, :r] = torch.randn(n_u, r, generator=g, device=dev) / math.sqrt(r)
V = new_storage(n_i // G_, r); V[:, :r] = torch.randn(n_i // G_, r, generator=g, device=dev) / math.sqrt(r)
rb = None
if bounded:
    ub = tdist.shard_bounds(n_u, G_)
    rb0 = tdist.topk_row_bounds(U, V, r, k, False, 0, ub[0], ub[1], n_i // G_)
    rb = rb0.repeat(G_)[:n_u].contiguous()  # stand-in: same distribution for every slice
for _ in range(2):
    score_topk(U, V, r, k, False, 0, row_bound=rb)
torch.cuda.synchronize()
