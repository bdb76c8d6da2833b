"""Multi-GPU plumbing (one process per GPU, ``torch.distributed``; NCCL on GPUs, gloo in CPU tests).

The reference has no distribution at all; this is the new-build design of SURVEY 8(e):

* training  -- users are sharded across ranks.  With identity user features each rank exclusively owns
  its users' rows of ``W_u`` (no communication); the item-side gradient ``dE_i`` (and any user-side
  parameters shared by all users: side-feature rows, biases, ReLU weights) is summed with ONE
  all-reduce per epoch, stream-ordered between the item-major pass and the Adam step.  The update is
  elementwise in the summed gradient, so replicas stay bit-identical.
* top-k     -- items are sharded; every rank scores all users against its own item slab
  (``tmf_score_topk`` with ``item_offset``), the per-rank ``[n_users, k]`` lists are all-gathered and
  merged with the (score desc, item id asc) comparator (``tmf_topk_merge``).  A global top-k member is
  always in its slab's local top-k, so the merge is exact.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n, world_size):
    """Contiguous, balanced ``[lo, hi)`` ranges: the first ``n % world_size`` ranks get one extra unit."""
    base, extra = divmod(int(n), int(world_size))
    bounds = [0]
    for rk in range(world_size):
        bounds.append(bounds[-1] + base + (1 if rk < extra else 0))
    return bounds


def balanced_user_bounds(row_lengths, world_size):
    """Contiguous user ranges with (nearly) equal numbers of interactions per rank."""
    lens = np.asarray(row_lengths, dtype=np.int64)
    csum = np.concatenate([[0], np.cumsum(lens)])
    total = csum[-1]
    bounds = [0]
    for rk in range(1, world_size):
        bounds.append(int(np.searchsorted(csum, total * rk / world_size, side="left")))
    bounds.append(len(lens))
    return [int(max(b, a)) for a, b in zip([0] + bounds[:-1], bounds)]


class GradientSync:
    """All-reduce hooks used by ``TrainPlan`` for user-sharded data-parallel training.

    ``shared_user_rows``: first row of ``W_u`` (or ``W_r`` for ReLU) that is shared by all ranks
    (side-feature rows after the rank-local identity block); ``None`` = no shared rows.
    """

    def __init__(self, group=None, shared_user_rows=None):
        self.group = group
        self.shared_user_rows = shared_user_rows
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.bytes_reduced = 0

    def _allreduce(self, t):
        if self.world_size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            self.bytes_reduced += t.numel() * t.element_size()

    def sync_item_grad(self, dEi):
        self._allreduce(dEi)

    def _shared_slices(self, u):
        out = []
        s = self.shared_user_rows
        if u.kind == "linear":
            if s is not None:
                out.append(("W", s))
        elif u.kind == "biased":
            if s is not None:
                out.append(("W", s))
            out.append(("b", 0))
        else:
            out.append(("W", 0))
            out.append(("br", 0))
            if s is not None:
                out.append(("Wr", s))
        return out

    def sync_shared_grads(self, u, i):
        # the item tower's gradients are functions of the already-reduced dE_i: nothing to do there
        for key, start in self._shared_slices(u):
            g = u.grads[key]
            self._allreduce(g[start:] if start else g)

    def broadcast_params(self, u, i):
        """Make replicated parameters identical at the start of fit (rank 0 wins)."""
        if self.world_size == 1:
            return
        for w in i.trainables().values():
            dist.broadcast(w, src=0, group=self.group)
        tr = u.trainables()
        for key, start in self._shared_slices(u):
            w = tr[key]
            dist.broadcast(w[start:] if start else w, src=0, group=self.group)

    def mean_loss(self, ip):
        from . import _engine as eng
        if ip.loss == eng.KL:
            raise NotImplementedError("KL's global moments are not sharded; run KL on one GPU")
        local = ip.mean_loss()
        t = torch.tensor([0.0 if ip.n_pos == 0 else local * ip.n_pos, float(ip.n_pos)], dtype=torch.float64,
                         device=ip.vals.device)
        self._allreduce(t)
        return float(t[0] / t[1]) if float(t[1]) > 0 else float("nan")


def merge_topk_lists(idx_lists, score_lists, k):
    """Host-side reference of the merge order (used by CPU tests): numpy, (score desc, id asc)."""
    idx = np.concatenate(idx_lists, axis=1)
    sc = np.concatenate(score_lists, axis=1)
    order = np.lexsort((idx, -sc.astype(np.float64)), axis=1)[:, :k]
    return np.take_along_axis(idx, order, 1), np.take_along_axis(sc, order, 1)


def sharded_topk(U, V_local, r, k, clamp, item_offset, group=None):
    """Item-sharded exact top-k: local fused scoring, all-gather of the ``[n_users, k]`` lists, merge.
    ``U`` (all users) and ``V_local`` (this rank's item slab) are padded storages on this rank's GPU.
    Every rank returns the full merged ``(idx, score)``."""
    from .. import _abi
    from .matrix_factorization import score_topk
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    k_local = min(k, V_local.shape[0])
    idx, sc = score_topk(U, V_local, r, k_local, clamp, item_offset)
    if world == 1:
        return idx, sc
    if k_local < k:  # tiny slab: pad with entries that can never win
        pad_i = torch.full((idx.shape[0], k - k_local), 2 ** 31 - 1, dtype=torch.int32, device=idx.device)
        pad_s = torch.full((idx.shape[0], k - k_local), float("-inf"), dtype=torch.float32, device=idx.device)
        idx, sc = torch.cat([idx, pad_i], 1).contiguous(), torch.cat([sc, pad_s], 1).contiguous()
    n_u = idx.shape[0]
    all_idx = torch.empty(world, n_u, k, dtype=torch.int32, device=idx.device)
    all_sc = torch.empty(world, n_u, k, dtype=torch.float32, device=idx.device)
    dist.all_gather_into_tensor(all_idx, idx, group=group)
    dist.all_gather_into_tensor(all_sc, sc, group=group)
    out_i = torch.empty(n_u, k, dtype=torch.int32, device=idx.device)
    out_s = torch.empty(n_u, k, dtype=torch.float32, device=idx.device)
    _abi.call("tmf_topk_merge", _abi.ptr(all_idx), _abi.ptr(all_sc), world, n_u, k, _abi.ptr(out_i), _abi.ptr(out_s))
    return out_i, out_s


def user_sharded_topk(U_local, V, r, k, clamp, group=None, gather=True):
    """Alternative to ``sharded_topk``: users are sharded, every rank holds all items.  No merge is needed (a
    row's top-k is computed entirely on one rank); with ``gather`` the ``[n_users, k]`` result is assembled on
    every rank by one all-gather (ranks must hold equally sized user slices, the last one may be padded)."""
    from .matrix_factorization import score_topk
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    idx, sc = score_topk(U_local, V, r, min(k, V.shape[0]), clamp, 0)
    if world == 1 or not gather:
        return idx, sc
    n_loc = idx.shape[0]
    all_idx = torch.empty(world * n_loc, idx.shape[1], dtype=torch.int32, device=idx.device)
    all_sc = torch.empty(world * n_loc, idx.shape[1], dtype=torch.float32, device=idx.device)
    dist.all_gather_into_tensor(all_idx, idx.contiguous(), group=group)
    dist.all_gather_into_tensor(all_sc, sc.contiguous(), group=group)
    return all_idx, all_sc
