"""Multi-GPU correctness check (run under torchrun on >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py

* user-sharded WMRB training step with [I_local | M] side features: the all-reduced item gradient, the
  reduced shared side-feature gradient and the rank-local user gradients must equal the single-GPU step;
* item-sharded fused top-k + all-gather + merge must equal the single-GPU top-k bit for bit.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
from scipy import sparse

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from teamoflow_b200.mf import dist as tdist  # noqa: E402
from teamoflow_b200.mf._engine import new_storage  # noqa: E402
from teamoflow_b200.mf._tensors import SparseInteractions  # noqa: E402
from teamoflow_b200.mf.initializer_graphs import Initializer  # noqa: E402
from teamoflow_b200.mf.loss_graphs import WMRBLoss  # noqa: E402
from teamoflow_b200.mf.matrix_factorization import MatrixFactorization, score_topk  # noqa: E402


class Fixed(Initializer):
    def __init__(self, W):
        self.W = W

    def initialize_weights(self, n_features, n_components):
        assert self.W.shape == (n_features, n_components), (self.W.shape, n_features, n_components)
        return torch.as_tensor(self.W, device="cuda")


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rng = np.random.default_rng(0)
    n_u, n_i, r, S, nnz, F = 4000, 3000, 64, 64, 200_000, 32
    cells = np.sort(rng.choice(n_u * n_i, nnz, replace=False))
    rows, cols = cells // n_i, cells % n_i
    vals = np.ones(nnz, np.float32)
    samp = np.stack([rng.choice(n_i, S, replace=False) for _ in range(n_u)]).astype(np.int64)
    M = sparse.random(n_u, F, density=4 / F, random_state=1, format="csr", dtype=np.float32)
    M.data[:] = 1
    Wu_id = (rng.random((n_u, r)) / 50).astype(np.float32)
    Wu_sh = (rng.random((F, r)) / 50).astype(np.float32)
    Wi = (rng.random((n_i, r)) / 50).astype(np.float32)

    def make(lo, hi, comm):
        Xu = sparse.hstack([sparse.eye(hi - lo, dtype=np.float32), M[lo:hi]]).tocsr()
        m = (rows >= lo) & (rows < hi)
        model = MatrixFactorization(r, loss_graph=WMRBLoss(), user_weight_graph=Fixed(np.concatenate([Wu_id[lo:hi], Wu_sh])),
                                    item_weight_graph=Fixed(Wi), n_users=hi - lo, n_items=n_i, n_samples=S)
        model.random_ind = torch.as_tensor(samp[lo:hi], device="cuda")
        inter = SparseInteractions(np.stack([rows[m] - lo, cols[m]], 1), vals[m], (hi - lo, n_i))
        from teamoflow_b200.mf._tensors import FeatureMatrix
        plan = model._prepare(Xu, FeatureMatrix.eye(n_i), inter, comm=comm)
        return plan

    b = tdist.balanced_user_bounds(np.bincount(rows, minlength=n_u), world)
    lo, hi = b[rank], b[rank + 1]
    ref = make(0, n_u, None)  # every rank also runs the whole problem on its own GPU
    ref.forward_backward()
    ok = True
    for peer in (True, False):  # NVLink peer-memory reduction (tmf_peer_reduce_push), then the NCCL all-reduce
        comm = tdist.GradientSync(shared_user_rows=hi - lo, peer=peer)
        plan = make(lo, hi, comm)
        comm.broadcast_params(plan.u, plan.i)
        plan.forward_backward()
        e_item = rel_err(plan.i.grads["W"], ref.i.grads["W"])
        e_user = rel_err(plan.u.grads["W"][:hi - lo], ref.u.grads["W"][lo:hi])
        e_shared = rel_err(plan.u.grads["W"][hi - lo:], ref.u.grads["W"][n_u:])
        loss_dp = comm.mean_loss(plan.ip)
        loss_1 = ref.ip.mean_loss()
        good = e_item < 1e-5 and e_user < 1e-5 and e_shared < 1e-5 and abs(loss_dp - loss_1) < 1e-5 * abs(loss_1)
        print(f"[rank {rank}] train ({'peer' if comm.peer else 'nccl'}{'' if comm.peer == peer else ' FALLBACK'}): rel err item {e_item:.2e} "
              f"user {e_user:.2e} shared {e_shared:.2e} loss {loss_dp:.6f}/{loss_1:.6f}", flush=True)
        ok = ok and good
        if comm.peer:
            # fused step: reduce-scatter + Adam step 1 + all-gather in one kernel == tmf_adam1 on the summed gradient, bit for bit,
            # and every replica holds the same bits
            from teamoflow_b200 import _abi
            want = plan.i.W.clone()
            g_full = plan.i.grads["W"].clone()
            _abi.call("tmf_adam1", _abi.ptr(want), _abi.ptr(g_full), want.numel(), 0.1)
            fused = plan.forward_backward(lr=0.1)
            same = bool(torch.equal(plan.i.W, want))
            allW = [torch.empty_like(want) for _ in range(world)]
            dist.all_gather(allW, plan.i.W.clone())
            repl = all(bool(torch.equal(w, allW[0])) for w in allW)
            print(f"[rank {rank}] fused reduce+adam: fused={fused} bit-exact={same} replicas identical={repl}", flush=True)
            ok = ok and fused and same and repl

    # ---- item-sharded top-k
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    nu2, ni2, r2, k = 3000, 50_000, 128, 100
    U = new_storage(nu2, r2); U[:, :r2] = torch.randn(nu2, r2, generator=g, device="cuda") / r2 ** 0.5
    V = new_storage(ni2, r2); V[:, :r2] = torch.randn(ni2, r2, generator=g, device="cuda") / r2 ** 0.5
    V[ni2 - 1] = V[0]  # a tie across shards
    ib = tdist.shard_bounds(ni2, world)
    for clamp in (False, True):
        idx1, sc1 = score_topk(U, V, r2, k, clamp)
        for exchange, bound in (("peer", "force"), ("peer", False), ("nccl", "force"), ("nccl", False)):
            try:
                idx, sc = tdist.sharded_topk(U, V[ib[rank]:ib[rank + 1]].contiguous(), r2, k, clamp, ib[rank], exchange=exchange, bound=bound)
                same = bool(torch.equal(idx, idx1) and torch.equal(sc, sc1))
                msg = "exact" if same else "MISMATCH"
            except RuntimeError as e:
                same, msg = False, f"ERROR {e}"
            print(f"[rank {rank}] sharded top-k clamp={clamp} exchange={exchange} bound={bound}: {msg}", flush=True)
            ok = ok and same
    # ---- ragged shapes: user count not divisible by the ranks, slabs smaller than k (padded lists), grid-valued ties
    for nu3, ni3, r3, k3 in ((1001, 40 * world + 3, 16, 50), (257, 997, 24, 20)):
        g3 = torch.Generator(device="cuda"); g3.manual_seed(6)
        U3 = new_storage(nu3, r3); U3[:, :r3] = torch.randint(-4, 5, (nu3, r3), generator=g3, device="cuda").float() / 8
        V3 = new_storage(ni3, r3); V3[:, :r3] = torch.randint(-4, 5, (ni3, r3), generator=g3, device="cuda").float() / 8
        ib3 = tdist.shard_bounds(ni3, world)
        for clamp in (False, True):
            idx1, sc1 = score_topk(U3, V3, r3, k3, clamp)
            for exchange, bound in (("peer", "force"), ("nccl", "force"), ("auto", True)):
                idx, sc = tdist.sharded_topk(U3, V3[ib3[rank]:ib3[rank + 1]].contiguous(), r3, k3, clamp, ib3[rank], exchange=exchange, bound=bound)
                same = bool(torch.equal(idx, idx1) and torch.equal(sc, sc1))
                print(f"[rank {rank}] ragged {nu3}x{ni3} k={k3} clamp={clamp} exchange={exchange} bound={bound}: {'exact' if same else 'MISMATCH'}", flush=True)
                ok = ok and same
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST CHECK OK" if int(flag) == 1 else "DIST CHECK FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
