"""1-GPU probe of the C4 strong-scaling split: builds each of the WORLD ranks' user range on ONE GPU in turn and times the
rank-local phases (user pass, item pass), to calibrate the cost model behind dist.balanced_user_bounds.

    python scripts/c4_balance_probe.py [world=8] [workload=c4]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
name = sys.argv[2] if len(sys.argv) > 2 else "c4"
torch.cuda.set_device(0)
for rank in range(world):
    wl = bench.Workload(name, rank, world, host_copy=False)
    xu, xi = wl.feature_args()
    plan = wl.model._prepare(xu, xi, wl.interactions(), comm=None)
    for _ in range(2):
        plan.step(wl.lr)
    ph = bench.profile_phases(plan, wl.lr, reps=3)
    print(f"rank {rank}: users {wl.w['n_u']} nnz {wl.nnz} user_pass {ph['user_pass']:.2f} ms item_pass {ph['item_pass']:.2f} ms "
          f"adam {ph['adam']:.2f} embed {ph['embed_fwd'] + ph['embed_bwd']:.2f}", flush=True)
    del plan, wl
    torch.cuda.empty_cache()
