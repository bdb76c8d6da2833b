"""Single-GPU probe of the item-sharded top-k protocol: plays rank 0 of a G-GPU run (its bound pass, its bounded
main pass over one slab, the peer merge of its user slice with local stand-in lists) and times each phase.
Usage: python scripts/topk_shard_probe.py [n_users n_items r k]"""
import ctypes as C
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from teamoflow_b200 import _abi  # noqa: E402
from teamoflow_b200.mf import dist as tdist  # noqa: E402
from teamoflow_b200.mf._engine import new_storage  # noqa: E402
from teamoflow_b200.mf.matrix_factorization import score_topk  # noqa: E402


def timed(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), out


def main():
    n_u, n_i, r, k = (int(x) for x in sys.argv[1:5]) if len(sys.argv) >= 5 else (1_000_000, 1_000_000, 128, 100)
    dev = torch.device("cuda")
    g = torch.Generator(device=dev); g.manual_seed(20245)
    U = new_storage(n_u, r); U[:, :r] = torch.randn(n_u, r, generator=g, device=dev) / math.sqrt(r)
    res = {"n_u": n_u, "n_i": n_i, "r": r, "k": k, "runs": []}
    for G_ in (8, 4, 2):
        n_slab = n_i // G_
        V = new_storage(n_slab, r); V[:, :r] = torch.randn(n_slab, r, generator=g, device=dev) / math.sqrt(r)
        ub = tdist.shard_bounds(n_u, G_)
        t_plain, (idx0, sc0) = timed(lambda: score_topk(U, V, r, k, False, 0))
        row = {"G": G_, "slab": n_slab, "main_unbounded_ms": t_plain, "variants": []}
        for n_s in sorted({tdist.bound_sample_size(n_slab, n_i, k), n_slab, max(n_slab // 4, 64 * k)}):
            t_b, b0 = timed(lambda: tdist.topk_row_bounds(U, V, r, k, False, 0, ub[0], ub[1], n_s))
            # stand-in bounds of all users: the same slab sample (the other ranks' samples are identically distributed)
            rb = tdist.topk_row_bounds(U, V, r, k, False, 0, 0, n_u, n_s)
            t_m, (idx1, sc1) = timed(lambda: score_topk(U, V, r, k, False, 0, row_bound=rb))
            real = float((idx1 != 2 ** 31 - 1).float().sum(1).mean())
            # exactness of the slab's contribution: every entry of the unbounded list that clears the bound must be present
            keep = sc0 >= rb[:, None]
            same = bool(((idx1 == idx0) | ~keep).all()) and bool(((idx1 == 2 ** 31 - 1) | keep).all())
            ptr_i = (C.c_void_p * G_)(*[idx1.data_ptr()] * G_)
            ptr_s = (C.c_void_p * G_)(*[sc1.data_ptr()] * G_)
            outs_i = [torch.empty(n_u, k, dtype=torch.int32, device=dev) for _ in range(2)]
            outs_s = [torch.empty(n_u, k, dtype=torch.float32, device=dev) for _ in range(2)]
            po_i = (C.c_void_p * G_)(*[outs_i[q % 2].data_ptr() for q in range(G_)])
            po_s = (C.c_void_p * G_)(*[outs_s[q % 2].data_ptr() for q in range(G_)])
            t_mg, _ = timed(lambda: _abi.call("tmf_topk_merge_peer", ptr_i, ptr_s, G_, ub[0], ub[1] - ub[0], k, po_i, po_s, G_))
            row["variants"].append({"n_s": n_s, "bound_ms": t_b, "main_bounded_ms": t_m, "merge_slice_ms": t_mg,
                                    "real_candidates_per_row": real, "consistent_with_unbounded": same,
                                    "rank_total_ms": t_b + t_m + t_mg})
            del rb, idx1, sc1, outs_i, outs_s
        res["runs"].append(row)
        print(json.dumps(row), flush=True)
        del V, idx0, sc0
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/topk_shard_probe.json", "w"), indent=1)


if __name__ == "__main__":
    main()
