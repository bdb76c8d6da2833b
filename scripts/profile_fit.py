"""Where does MatrixFactorization.fit() spend its set-up time?  cProfile of one fit(1) from host buffers with
CUDA_LAUNCH_BLOCKING=1 (every launch blocks, so cumulative times include the GPU work each Python call starts).
Development aid; run on a GPU box:  CUDA_LAUNCH_BLOCKING=1 python scripts/profile_fit.py [workload]"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
torch.cuda.set_device(0)
wl = bench.Workload(name, 0, 1)
xu, xi = wl.feature_args()
for _ in range(2):
    wl.model.fit(1, xu, xi, wl.interactions(), lr=wl.lr, verbose=False)
torch.cuda.synchronize()
t0 = time.perf_counter()
wl.model.fit(1, xu, xi, wl.interactions(), lr=wl.lr, verbose=False)
torch.cuda.synchronize()
print("fit(1) wall: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
pr = cProfile.Profile()
pr.enable()
wl.model.fit(1, xu, xi, wl.interactions(), lr=wl.lr, verbose=False)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
