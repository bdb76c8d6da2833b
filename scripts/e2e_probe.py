"""Where does MatrixFactorization.fit(20) from host buffers spend its time?  Wall-clock of _prepare / the epoch loop / the
graph capture, synchronised, three calls in a row like bench.py's e2e leg.  Development aid (GPU box)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from teamoflow_b200.mf import _engine as eng  # noqa: E402
from teamoflow_b200.mf.matrix_factorization import MatrixFactorization  # noqa: E402

torch.cuda.set_device(0)
wl = bench.Workload(sys.argv[1] if len(sys.argv) > 1 else "c3", 0, 1)
xu, xi = wl.feature_args()
T = {}


def timed(name, fn):
    def wrapper(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = fn(*a, **k)
        torch.cuda.synchronize(); T[name] = T.get(name, 0.0) + time.perf_counter() - t0
        return out
    return wrapper


MatrixFactorization._prepare = timed("prepare", MatrixFactorization._prepare)
eng.TrainPlan.run = timed("run", eng.TrainPlan.run)
eng.TrainPlan._captured_step = timed("capture", eng.TrainPlan._captured_step)
eng.InteractionPlan.__init__ = timed("  InteractionPlan", eng.InteractionPlan.__init__)
eng.InteractionPlan.set_samples = timed("    set_samples", eng.InteractionPlan.set_samples)
eng.InteractionPlan._build_work_list = timed("    work_list", eng.InteractionPlan._build_work_list)
MatrixFactorization._make_tower = timed("  make_tower", MatrixFactorization._make_tower)
import teamoflow_b200.mf.matrix_factorization as mfm  # noqa: E402
mfm.as_interactions = timed("  as_interactions(H2D)", mfm.as_interactions)
mfm.as_features = timed("  as_features", mfm.as_features)
for i in range(4):
    T.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    wl.model.fit(20, xu, xi, wl.interactions(), lr=wl.lr, verbose=False)
    loss = wl.model._plan.ip.mean_loss()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"fit(20) call {i}: {dt * 1e3:.1f} ms | " + " | ".join(f"{k.strip()} {v * 1e3:.1f}" for k, v in T.items()))
