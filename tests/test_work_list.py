"""Host logic of the user pass's work list (teamoflow_b200.mf._engine.InteractionPlan): every interaction is covered by
exactly one work item, no item is longer than the slice length, heavy users come first, and the slot bookkeeping the
fix-up kernel relies on is consistent.  Runs on CPU tensors (no kernel is launched)."""
import numpy as np
import torch

from teamoflow_b200.mf import _engine as eng


def _plan(lens):
    p = object.__new__(eng.InteractionPlan)
    lens = np.asarray(lens, dtype=np.int64)
    p.n_users = int(lens.size)
    p.nnz = int(lens.sum())
    p.row_ptr = torch.as_tensor(np.concatenate([[0], np.cumsum(lens)]).astype(np.int32))
    p.vals = torch.ones(max(p.nnz, 1))
    p.S = 0
    return p


def test_slice_length_adapts_to_the_problem_size():
    p = _plan([10])
    for nnz, want in ((10_000, eng.InteractionPlan.SPLIT_MIN), (56_000, eng.InteractionPlan.SPLIT_MIN),
                      (20_000_000, eng.InteractionPlan.SPLIT), (500_000_000, eng.InteractionPlan.SPLIT)):
        p.nnz = nnz
        assert p.slice_len() == want
    p.nnz = 4_736_000  # in between: nnz / (2 x resident warp slots)
    assert eng.InteractionPlan.SPLIT_MIN < p.slice_len() < eng.InteractionPlan.SPLIT


def test_work_list_covers_every_interaction_once_and_respects_the_slice_length():
    rng = np.random.default_rng(3)
    lens = np.concatenate([[0, 1, 5000, 33, 32, 31, 2049, 0], rng.integers(0, 200, 300)])
    p = _plan(lens)
    p._build_work_list()
    L = p.slice_len()
    wu, wa, wb, ws = (t.numpy().astype(np.int64) for t in (p.w_user, p.w_a, p.w_b, p.w_slot))
    assert p.n_work == wu.size and ((wb - wa) <= L).all() and ((wb - wa) >= 0).all()
    assert (np.diff(wb - wa) <= 0).all()  # heaviest first
    rp = p.row_ptr.numpy().astype(np.int64)
    assert (wa >= rp[wu]).all() and (wb <= rp[wu + 1]).all()
    cover = np.zeros(p.nnz, dtype=np.int64)
    for a, b in zip(wa, wb):
        cover[a:b] += 1
    assert (cover == 1).all()
    # users with more than L interactions are split; their slices own consecutive slots starting at split_first
    split_users = np.nonzero(lens > L)[0]
    assert p.n_split == split_users.size and np.array_equal(p.split_user.numpy(), split_users)
    nseg = p.split_nseg.numpy().astype(np.int64)
    assert np.array_equal(nseg, -(-lens[split_users] // L)) and p.n_slots == nseg.sum()
    first = p.split_first.numpy().astype(np.int64)
    assert np.array_equal(first, np.concatenate([[0], np.cumsum(nseg)[:-1]]))
    for u, f, n in zip(split_users, first, nseg):
        slots = np.sort(ws[wu == u])
        assert np.array_equal(slots, np.arange(f, f + n))
    assert (ws[~np.isin(wu, split_users)] == -1).all()
