#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/tk.py <<'PY'
import torch, math, sys, os
sys.path.insert(0, '.')
from teamoflow_b200 import _abi
from teamoflow_b200.mf._engine import new_storage
from teamoflow_b200.mf.matrix_factorization import score_topk
for (n_u, n_i) in ((151552, 1000000),):
    r, k = 128, 100
    g = torch.Generator(device='cuda'); g.manual_seed(1)
    U = new_storage(n_u, r); U[:, :r] = torch.randn(n_u, r, generator=g, device='cuda') / math.sqrt(r)
    V = new_storage(n_i, r); V[:, :r] = torch.randn(n_i, r, generator=g, device='cuda') / math.sqrt(r)
    for it in range(int(os.environ.get('REPS', '3'))):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); score_topk(U, V, r, k, False); e1.record(); torch.cuda.synchronize()
    print('dbg', os.environ.get('TMF_TOPK_DEBUG'), n_u, n_i, 'total ms', e0.elapsed_time(e1), 'TF/s', 2*n_u*n_i*r/e0.elapsed_time(e1)/1e9, flush=True)
PY
timeout 900 python -m pytest tests/test_gpu_score.py -m gpu -q --timeout 300 -x 2>&1 | tail -5
timeout 300 python /tmp/tk.py 2>&1 | tee gpurun_out/topk_time.log && \
REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_topk_v5.csv python /tmp/tk.py > gpurun_out/ncu_topk.log 2>&1
