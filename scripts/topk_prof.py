"""Cycle-counter breakdown of the fused top-k kernel (TMF_TOPK_PROF=1): python scripts/topk_prof.py n_users n_items"""
import math, os, sys
os.environ["TMF_TOPK_PROF"] = "1"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from teamoflow_b200.mf._engine import new_storage
from teamoflow_b200.mf.matrix_factorization import score_topk
n_u, n_i = int(sys.argv[1]), int(sys.argv[2]); r, k = 128, 100
dev = torch.device("cuda")
g = torch.Generator(device=dev); g.manual_seed(20245)
U = new_storage(n_u, r); U[:, :r] = torch.randn(n_u, r, generator=g, device=dev) / math.sqrt(r)
V = new_storage(n_i, r); V[:, :r] = torch.randn(n_i, r, generator=g, device=dev) / math.sqrt(r)
for _ in range(2):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); score_topk(U, V, r, k, False, 0); e1.record(); torch.cuda.synchronize()
    print("ms", e0.elapsed_time(e1), flush=True)
