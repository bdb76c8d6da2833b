"""Multi-GPU plumbing (one process per GPU, ``torch.distributed``; NCCL on GPUs, gloo in CPU tests).

The reference has no distribution at all; this is the new-build design of SURVEY 8(e):

* training  -- users are sharded across ranks.  With identity user features each rank exclusively owns
  its users' rows of ``W_u`` (no communication); the item-side gradient ``dE_i`` (and any user-side
  parameters shared by all users: side-feature rows, biases, ReLU weights) is summed with ONE
  all-reduce per epoch, stream-ordered between the item-major pass and the Adam step.  The update is
  elementwise in the summed gradient, so replicas stay bit-identical.
* top-k     -- items are sharded; every rank scores all users against its own item slab
  (``tmf_score_topk_bounded`` with ``item_offset`` and cross-rank score bounds), the per-rank ``[n_users, k]``
  lists are exchanged and merged with the (score desc, item id asc) comparator -- over NVLink peer memory in one
  kernel (``tmf_topk_merge_peer``) or through NCCL (``tmf_topk_merge``).  A global top-k member is always in
  its slab's local top-k, so the merge is exact.
"""
from __future__ import annotations

import os

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


def rebalanced_user_bounds(row_weights, bounds, rank_times):
    """Measure-and-rebalance pass for a user-sharded split.  ``bounds``: the current contiguous user ranges, ``rank_times``: the
    compute time each rank measured for its range (same unit for all), ``row_weights``: the per-user weights the first split
    used (interactions + sampled negatives).  What a unit of weight costs differs between the ranges -- the heaviest users'
    ranges gather from a user table that fits the L2, the lightest users' ranges pay a fixed cost per user -- so each range is
    taken to have its own constant cost per unit of weight, and the new boundaries sit at equal shares of the total cost."""
    w = np.asarray(row_weights, dtype=np.float64)
    world = len(bounds) - 1
    cost = np.empty_like(w)
    for rk in range(world):
        a, b = bounds[rk], bounds[rk + 1]
        if b > a:
            cost[a:b] = w[a:b] * (float(rank_times[rk]) / max(float(w[a:b].sum()), 1e-30))
    ccost = np.concatenate([[0.0], np.cumsum(cost)])
    new = [0]
    for rk in range(1, world):
        new.append(int(np.searchsorted(ccost, ccost[-1] * rk / world, side="left")))
    new.append(len(w))
    return [int(max(b, a)) for a, b in zip([0] + new[:-1], new)]


class GradientSync:
    """Gradient exchange used by ``TrainPlan`` for user-sharded data-parallel training.

    ``shared_user_rows``: first row of ``W_u`` (or ``W_r`` for ReLU) that is shared by all ranks
    (side-feature rows after the rank-local identity block); ``None`` = no shared rows.

    Peer path (default on GPUs): the item-side gradient lives in the NVLink peer arena and is summed by
    ``tmf_peer_reduce_push`` (Adam step fused in for an identity-feature Linear item tower); the small gradients of
    the shared user-side parameters are staged into the same arena and summed by a second launch between the SAME
    two flag barriers -- a training step holds no NCCL work, so its captured CUDA graph is plain kernel launches.
    Fallback (``peer=False`` or CUDA IPC refused): NCCL all-reduces.
    """

    def __init__(self, group=None, shared_user_rows=None, peer=True):
        self.group = group
        self.shared_user_rows = shared_user_rows
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.bytes_reduced = 0
        self.peer = bool(peer) and self.world_size > 1 and torch.cuda.is_available()
        self._ar = None
        self._plan = None
        self._dE = self._W = self._stage = None
        self._dE_off = self._W_off = self._stage_off = None
        self._rows = None

    def _allreduce(self, t):
        if self.world_size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            self.bytes_reduced += t.numel() * t.element_size()

    # ---- peer arena life cycle

    def attach(self, plan):
        """Move the item tower's gradient buffer (and, when the update can be fused, its weights) into peer memory and
        reserve the staging block of the shared user-side gradients.  Collective; called by ``TrainPlan``.  The arena
        is cached per group: one attached plan at a time."""
        if not self.peer:
            return
        it = plan.i
        a256 = lambda n: (int(n) + 255) // 256 * 256  # noqa: E731
        fuse = it.kind == "linear" and it.X.identity
        tr = plan.u.trainables()
        n_sh = sum(int((tr[k][start:] if start else tr[k]).numel()) for k, start in self._shared_slices(plan.u))
        n_sh4 = (n_sh + 3) // 4  # float4 units, zero padded
        n_dE, n_W, n_st = a256(it.dE.numel() * 4), (a256(it.W.numel() * 4) if fuse else 0), a256(n_sh4 * 16)
        ar = peer_arena(n_dE + n_W + n_st, self.group, tag="grad")
        if ar is None:
            self.peer = False
            return
        for other in list(ar.views):  # a plan attached earlier (previous fit) gives the arena up first
            if other is not self or other._plan is not plan:
                other._arena_closing(ar)
        self._ar, self._plan = ar, plan
        ar.views = [self]
        self._dE_off, self._W_off, self._stage_off = 0, (n_dE if fuse else None), n_dE + n_W
        self._dE = ar.local(0, tuple(it.dE.shape), torch.float32)
        self._dE.zero_()
        it.dE = self._dE
        if fuse:
            self._W = ar.local(n_dE, tuple(it.W.shape), torch.float32)
            self._W.copy_(it.W)
            it.W = it.E = self._W
        self._stage = None
        if n_sh4:
            self._stage = ar.local(self._stage_off, (n_sh4 * 4,), torch.float32)
            self._stage.zero_()
            self._stage_bounds = shard_bounds(n_sh4, self.world_size)
        self._rows = shard_bounds(it.dE.shape[0], self.world_size)

    def _arena_closing(self, arena):
        """The arena is about to be freed or handed to another plan: tensors of the attached plan move out of it."""
        plan = self._plan
        if plan is not None and self._ar is arena:
            it = plan.i
            if self._dE is not None and it.dE.data_ptr() == self._dE.data_ptr():
                it.dE = it.dE.clone()
            if self._W is not None and it.W.data_ptr() == self._W.data_ptr():
                alias = it.E.data_ptr() == it.W.data_ptr()
                it.W = it.W.clone()
                if alias:
                    it.E = it.W
            if hasattr(plan, "invalidate_graph"):
                plan.invalidate_graph()  # a captured step addresses the arena
        if arena is not None and self in arena.views:
            arena.views.remove(self)
        self._ar = self._plan = None
        self._dE = self._W = self._stage = None

    def detach(self, plan):
        """End of fit: the plan's tensors leave the (shared, reusable) arena; barrier time-outs surface here."""
        ar = self._ar
        if ar is not None:
            ar.check()
        if self._plan is plan:
            self._arena_closing(ar)

    # ---- the exchange of one step

    def sync_grads(self, plan, lr=None):
        """Everything a step exchanges, after the item-major pass and the USER tower's backward: ``dE_i`` (summed; with
        ``lr`` the Adam step of an identity Linear item tower is fused in -> returns True and the caller skips that
        update) and the shared user-side gradients (summed in place)."""
        if self.world_size == 1:
            return False
        slices = self._shared_slices(plan.u)
        peer = self._ar is not None and self._dE is not None and plan.i.dE.data_ptr() == self._dE.data_ptr()
        if not peer:
            self._allreduce(plan.i.dE)
            self.sync_shared_grads(plan.u, plan.i)
            return False
        views = []
        if self._stage is not None:
            off = 0
            for key, start in slices:
                g = plan.u.grads[key]
                g = (g[start:] if start else g).reshape(-1)
                self._stage[off:off + g.numel()].copy_(g)
                views.append((g, off))
                off += g.numel()
        fused = self.sync_item_grad(plan.i.dE, lr=lr, tower=plan.i, _with_stage=True)
        for g, off in views:
            g.copy_(self._stage[off:off + g.numel()])
        return fused

    def sync_item_grad(self, dEi, lr=None, tower=None, _with_stage=False):
        """Sum ``dE_i`` over the ranks (in place).  Returns True when the Adam step of ``tower.W`` was fused in
        (then ``dEi`` keeps the LOCAL partial and the caller must skip that update)."""
        if self.world_size == 1:
            return False
        if self._ar is None or self._dE is None or dEi.data_ptr() != self._dE.data_ptr():
            self._allreduce(dEi)
            return False
        from .. import _abi
        ar, ld = self._ar, dEi.shape[1]
        lo, hi = self._rows[self.rank], self._rows[self.rank + 1]
        fused = (lr is not None and self._W is not None and tower is not None and tower.W.data_ptr() == self._W.data_ptr())
        ar.barrier()  # every rank's partials (and staged shared gradients) are complete
        _abi.call("tmf_peer_reduce_push", ar.ptrs(self._dE_off), ar.ptrs(self._W_off if fused else self._dE_off), self.world_size,
                  self.rank, lo * ld, (hi - lo) * ld, float(lr) if fused else -1.0)
        self.bytes_reduced += dEi.numel() * 4
        if _with_stage and self._stage is not None:
            a, b = self._stage_bounds[self.rank], self._stage_bounds[self.rank + 1]
            if b > a:
                _abi.call("tmf_peer_reduce_push", ar.ptrs(self._stage_off), ar.ptrs(self._stage_off), self.world_size, self.rank,
                          4 * a, 4 * (b - a), -1.0)
            self.bytes_reduced += self._stage.numel() * 4
        ar.barrier()  # every rank's slices have landed everywhere; partial buffers may be rewritten
        return fused

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
        """NCCL form of the shared user-side gradient exchange (the peer path stages them, see ``sync_grads``)."""
        # the item tower's gradients are functions of the already-reduced dE_i: nothing to do there
        for key, start in self._shared_slices(u):
            g = u.grads[key]
            self._allreduce(g[start:] if start else g)

    def allreduce_kl_moments(self, moments):
        """KL under user sharding: the six additive fp64 sums of ``tmf_kl_moments`` (48 bytes) summed over the ranks."""
        self._allreduce(moments)

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
        if self._ar is not None:
            self._ar.check()
        if ip.loss == eng.KL:
            return ip.mean_loss()  # the scalar is already global: every rank derived it from the summed moments
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


# ----------------------------------------------------------------------------- NVLink peer memory


class _DevMem:
    """Raw device memory as a ``__cuda_array_interface__`` provider (zero-copy ``torch.as_tensor``)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class PeerArena:
    """A block of this rank's HBM that every rank of the group can address directly over NVLink (CUDA IPC
    mappings of plain ``cudaMalloc`` memory, ``tmf_peer_alloc`` / ``tmf_ipc_*``), plus the flag pad of the
    stream-ordered barrier (``tmf_peer_barrier``).  Collective: every rank constructs it with the same size.

    Construction is collective-safe: the ranks agree (all-reduce MIN) after the local allocation + export and again
    after mapping the peers' handles, so a rank that fails (out of memory, IPC refused) never leaves the others
    inside a different collective; whatever the ranks that succeeded had built is released again and ``PeerArena.create``
    returns ``None`` on every rank."""

    PAD_BYTES = 256
    ERR_WORD = 33  # uint32 index inside the pad: 1 + missing rank after a barrier timeout (csrc/peer.cu kErrWord)

    def __init__(self):
        raise TypeError("use PeerArena.create(nbytes, group)")

    @classmethod
    def create(cls, nbytes, group=None):
        import ctypes as C
        import sys
        from .. import _abi
        self = object.__new__(cls)
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.nbytes = (int(nbytes) + 255) // 256 * 256
        self.closed = False
        self.views = []  # GradientSync instances whose plans hold tensors inside this arena (told before it is freed)
        self._base = None
        self.bases = []
        self.timeout_ms = int(float(os.environ.get("TMF_PEER_TIMEOUT_S", "0")) * 1000)  # 0 = library default (120 s)
        dev = torch.device("cuda", torch.cuda.current_device())

        def agree(ok):
            flag = torch.tensor([1 if ok else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            return int(flag) == 1

        def report(stage, e):
            print(f"[teamoflow_b200] peer memory unavailable on rank {self.rank} ({stage}: {type(e).__name__}: {e}); "
                  "falling back to the NCCL exchange", file=sys.stderr, flush=True)

        # ---- stage 1: local allocation + export
        handle, ok = C.create_string_buffer(64), True
        try:
            base = C.c_void_p()
            _abi.call_nostream("tmf_peer_alloc", self.PAD_BYTES + self.nbytes, C.byref(base))
            self._base = base.value
            _abi.call_nostream("tmf_ipc_export", C.c_void_p(self._base), handle)
        except Exception as e:  # noqa: BLE001
            ok = False
            report("allocate/export", e)
        if not agree(ok):
            self._release()
            return None
        # ---- stage 2: exchange the handles (every rank reaches this point) and map the peers' arenas
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=group)
        self.bases = [None] * self.world
        try:
            for g, h in enumerate(handles):
                if g == self.rank:
                    self.bases[g] = self._base
                else:
                    ptr = C.c_void_p()
                    _abi.call_nostream("tmf_ipc_open", C.create_string_buffer(h, 64), C.byref(ptr))
                    self.bases[g] = ptr.value
        except Exception as e:  # noqa: BLE001
            ok = False
            report("map peers", e)
        if not agree(ok):
            self._release()
            return None
        self._pad = torch.as_tensor(_DevMem(self._base, self.PAD_BYTES), device=dev).view(torch.int32)
        self._mem = torch.as_tensor(_DevMem(self._base + self.PAD_BYTES, self.nbytes), device=dev)
        self._pads = (C.c_void_p * self.world)(*self.bases)
        return self

    def _release(self):
        """Undo whatever this rank built (no collective inside)."""
        import ctypes as C
        from .. import _abi
        for g, b in enumerate(self.bases):
            if g != self.rank and b is not None:
                try:
                    _abi.call_nostream("tmf_ipc_close", C.c_void_p(b))
                except Exception:  # noqa: BLE001
                    pass
        self.bases = []
        self._mem = self._pad = None
        if self._base is not None:
            try:
                _abi.call_nostream("tmf_peer_free", C.c_void_p(self._base))
            except Exception:  # noqa: BLE001
                pass
            self._base = None
        self.closed = True

    def local(self, offset, shape, dtype):
        """Tensor view of this rank's arena at byte ``offset`` (256-byte aligned offsets)."""
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        assert not self.closed and offset % 256 == 0 and offset + n <= self.nbytes
        return self._mem[offset:offset + n].view(dtype).reshape(shape)

    def ptrs(self, offset):
        """Host array of the ``world`` device pointers to byte ``offset`` of every rank's arena."""
        import ctypes as C
        return (C.c_void_p * self.world)(*[b + self.PAD_BYTES + int(offset) for b in self.bases])

    def barrier(self):
        """Stream-ordered barrier over the group (no host synchronisation)."""
        from .. import _abi
        # epoch 0 = "next value of the device-side counter in this rank's pad": no per-call argument, so the launch can
        # be replayed from a CUDA graph (TrainPlan.run)
        _abi.call("tmf_peer_barrier", self._pads, self.world, self.rank, 0, self.timeout_ms)

    def check(self):
        """Raise if a barrier of this arena timed out on this rank (synchronises: reads one word of the pad).  Called
        where the host synchronises anyway (loss read-out, end of fit, after a timed top-k)."""
        if self.closed:
            return
        err = int(self._pad[self.ERR_WORD].item())
        if err:
            raise RuntimeError(f"peer barrier timed out on rank {self.rank} waiting for rank {err - 1} "
                               "(TMF_PEER_TIMEOUT_S sets the limit); results of this exchange are invalid")

    def close(self):
        """Collective: every rank's kernels have drained, then the mappings and the allocation go."""
        if self.closed:
            return
        for v in list(self.views):
            v._arena_closing(self)
        self.views = []
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        self._release()


_arenas = {}
_peer_disabled = {}


def peer_arena(nbytes, group=None, tag="default"):
    """The cached arena of ``group`` (grown collectively when too small), or None when peer memory is unavailable
    (IPC refused by the platform): callers then use the NCCL exchange.  The failure is reported once on stderr."""
    key = (id(group) if group is not None else 0, tag)
    if _peer_disabled.get(key):
        return None
    ar = _arenas.get(key)
    if ar is not None and ar.nbytes >= nbytes:
        return ar
    if ar is not None:
        ar.close()  # tensors of attached plans are moved out first (GradientSync._arena_closing)
        _arenas.pop(key)
    ar = PeerArena.create(nbytes, group)
    if ar is None:
        _peer_disabled[key] = True
        return None
    _arenas[key] = ar
    return ar


def exchange_mode(group=None, tag="topk"):
    """``"peer"`` when the group's arena of that purpose is live, else ``"nccl"`` (reporting aid)."""
    return "peer" if (id(group) if group is not None else 0, tag) in _arenas else "nccl"


# ----------------------------------------------------------------------------- item-sharded top-k


def bound_sample_size(n_local_items, n_items_total, k, world=None):
    """Items of the own slab the bound pass scores (0 = no bound pass).  The k-th best score against ``n_s`` items
    lets about ``k * n_local / n_s`` candidates per (user, slab) through the main pass and removes its per-row
    warm-up, for ``n_u/G * n_s`` extra score pairs.  Measured on B200 (1M x 1M, r=128, k=100, one rank's work,
    ``scripts/topk_shard_probe.py``): G=8 100 -> 66 ms and G=4 139 -> 113 ms with the whole slab as the sample,
    G=2 206 -> 225+ ms with any sample -- so the pass runs from 4 ranks up."""
    if world is not None and world < 4:
        return 0
    return int(n_local_items)


def topk_row_bounds(U, V_local, r, k, clamp, item_offset, lo_u, hi_u, n_s, with_lists=False):
    """Bound pass of one rank: exact k-th best canonical score of users ``[lo_u, hi_u)`` against the first ``n_s``
    items of the own slab -- a lower bound of those users' global k-th best score (``-inf`` when the slab is tiny).
    ``with_lists``: also return the ``(idx, score)`` lists the bounds were read from."""
    from .matrix_factorization import score_topk
    n = hi_u - lo_u
    if n_s < k or n == 0:
        b = torch.full((n,), float("-inf"), dtype=torch.float32, device=U.device)
        return (b, None, None) if with_lists else b
    idx, sc = score_topk(U[lo_u:hi_u], V_local[:n_s], r, k, clamp, item_offset)
    b = sc[:, k - 1].contiguous()
    return (b, idx, sc) if with_lists else b


def _gather_rows(local, bounds, group):
    """all-gather of row blocks of unequal sizes ``bounds[g+1]-bounds[g]`` (padded to the largest)."""
    world = len(bounds) - 1
    n_max = max(bounds[g + 1] - bounds[g] for g in range(world))
    pad = torch.zeros((n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    allp = torch.empty((world * n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(allp, pad, group=group)
    if all(bounds[g + 1] - bounds[g] == n_max for g in range(world)):
        return allp
    return torch.cat([allp[g * n_max:g * n_max + bounds[g + 1] - bounds[g]] for g in range(world)])


def sharded_topk(U, V_local, r, k, clamp, item_offset, group=None, exchange="auto", bound=True, n_items_total=None, events=None):
    """Item-sharded exact top-k (north_star: "each GPU scoring its own item slab ... merged by allgather").

    1. bound pass -- rank g scores ITS SLICE of the users against a sample of its slab; the k-th best score is a
       lower bound of those users' global k-th best; the bounds are all-gathered (4 bytes per user);
    2. every rank scores ALL users against its slab with those bounds (``tmf_score_topk_bounded``): per (user, slab)
       only candidates that can still reach the global top-k are listed and reranked, so the per-row work shrinks
       with the slab instead of being repeated on every GPU;
    3. exchange -- ``"peer"``: one kernel per rank pulls its user slice's G lists over NVLink, merges them and pushes
       the merged rows to every rank (``tmf_topk_merge_peer``, all-to-all + merge + all-gather in one launch);
       ``"nccl"``: all-to-all, ``tmf_topk_merge``, all-gather.  ``"auto"`` = peer when IPC works.
    ``U`` (all users) and ``V_local`` (this rank's slab) are padded storages.  Every rank returns the full merged
    ``(idx, score)``, identical to ``score_topk`` over the concatenated slabs.  ``events``: optional list that
    receives ``(phase name, CUDA event)`` marks (bench.py's phase breakdown)."""

    def mark(name):
        if events is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            events.append((name, e))

    from .. import _abi
    from .matrix_factorization import score_topk
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    n_loc_items = V_local.shape[0]
    if world == 1:
        return score_topk(U, V_local, r, min(k, n_loc_items), clamp, item_offset)
    rank = dist.get_rank(group)
    n_u = U.shape[0]
    dev = U.device
    ub = shard_bounds(n_u, world)
    lo_u, hi_u = ub[rank], ub[rank + 1]
    # ---- 1. bounds
    mark("start")
    row_bound = own = None
    if bound:
        n_s = bound_sample_size(n_loc_items, n_items_total, k, world) if bound != "force" else n_loc_items
        if n_s > 0:
            b_loc, own_i, own_s = topk_row_bounds(U, V_local, r, k, clamp, item_offset, lo_u, hi_u, n_s, with_lists=True)
            row_bound = _gather_rows(b_loc, ub, group).contiguous()
            if n_s == n_loc_items and own_i is not None:
                own = (own_i, own_s)  # the whole slab was scored: these ARE this rank's lists of its own user slice
    mark("bound_pass")
    # ---- 2. local lists (in peer memory when the peer exchange is used)
    arena = None
    if exchange in ("auto", "peer"):
        list_bytes = (n_u * k * 4 + 255) // 256 * 256
        arena = peer_arena(4 * list_bytes, group, tag="topk")
        if arena is None and exchange == "peer":
            raise RuntimeError("peer-memory exchange requested but CUDA IPC is unavailable")
    k_local = min(k, n_loc_items)
    if arena is not None:
        idx = arena.local(0, (n_u, k), torch.int32)
        sc = arena.local(list_bytes, (n_u, k), torch.float32)
    else:
        idx = torch.empty(n_u, k, dtype=torch.int32, device=dev)
        sc = torch.empty(n_u, k, dtype=torch.float32, device=dev)
    if k_local < k:  # tiny slab: pad with entries that can never win
        ti, ts = score_topk(U, V_local, r, k_local, clamp, item_offset, row_bound=row_bound)
        idx[:, :k_local], sc[:, :k_local] = ti, ts
        idx[:, k_local:], sc[:, k_local:] = 2 ** 31 - 1, float("-inf")
    elif own is not None:  # rows of the own slice come from the bound pass, the others are scored against the bounds
        idx[lo_u:hi_u], sc[lo_u:hi_u] = own
        for a, b in ((0, lo_u), (hi_u, n_u)):
            if b > a:
                score_topk(U[a:b], V_local, r, k, clamp, item_offset, row_bound=row_bound[a:b], out=(idx[a:b], sc[a:b]))
    else:
        score_topk(U, V_local, r, k, clamp, item_offset, row_bound=row_bound, out=(idx, sc))
    mark("slab_scoring")
    # ---- 3. exchange + merge
    if arena is not None:
        arena.barrier()  # every rank's lists are complete and visible
        _abi.call("tmf_topk_merge_peer", arena.ptrs(0), arena.ptrs(list_bytes), world, lo_u, hi_u - lo_u, k,
                  arena.ptrs(2 * list_bytes), arena.ptrs(3 * list_bytes), world)
        arena.barrier()  # every rank's pushes have landed; the lists may be overwritten by the next call
        res = (arena.local(2 * list_bytes, (n_u, k), torch.int32).clone(), arena.local(3 * list_bytes, (n_u, k), torch.float32).clone())
        mark("exchange_merge")
        return res
    n_loc = hi_u - lo_u
    splits = [ub[g + 1] - ub[g] for g in range(world)]
    r_idx = torch.empty(world * n_loc, k, dtype=torch.int32, device=dev)
    r_sc = torch.empty(world * n_loc, k, dtype=torch.float32, device=dev)
    dist.all_to_all_single(r_idx, idx, output_split_sizes=[n_loc] * world, input_split_sizes=splits, group=group)
    dist.all_to_all_single(r_sc, sc, output_split_sizes=[n_loc] * world, input_split_sizes=splits, group=group)
    m_idx = torch.empty(n_loc, k, dtype=torch.int32, device=dev)
    m_sc = torch.empty(n_loc, k, dtype=torch.float32, device=dev)
    _abi.call("tmf_topk_merge", _abi.ptr(r_idx), _abi.ptr(r_sc), world, n_loc, k, _abi.ptr(m_idx), _abi.ptr(m_sc))
    res = (_gather_rows(m_idx, ub, group), _gather_rows(m_sc, ub, group))
    mark("exchange_merge")
    return res


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
