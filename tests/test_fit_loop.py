"""Host logic of ``MatrixFactorization.fit``'s epoch loop (ref: matrix_factorization.py:128-183): the epochs are handed to
``TrainPlan.run`` in spans that end at every 25th epoch (loss report) and at every resampling point; the total number of
steps is ``epochs``; the captured step is dropped when the negatives change and when fit returns.  Runs on CPU with a
recording stand-in for the plan (no kernel is launched)."""
import numpy as np
import pytest
import torch

from teamoflow_b200.mf import _engine as eng
from teamoflow_b200.mf import matrix_factorization as mfm


class _Tower:
    kind = eng.LINEAR

    def __init__(self, n, r):
        self.W = torch.zeros(n, 4 * ((r + 3) // 4))

    def forward(self):
        return self.W


class _IP:
    loss = eng.WMRB
    n_items, n_users, S = 7, 5, 3

    def __init__(self, log):
        self.log = log

    def set_samples(self, ri):
        self.log.append(("resample", None))

    def mean_loss(self):
        self.log.append(("loss", None))
        return 0.5


class _Plan:
    opt_state = None

    def __init__(self, log):
        self.log = log
        self.u, self.i, self.ip = _Tower(5, 4), _Tower(7, 4), _IP(log)

    def run(self, n, lr):
        self.log.append(("run", int(n)))

    def invalidate_graph(self):
        self.log.append(("invalidate", None))


@pytest.mark.parametrize("epochs,resample,spans", [
    (7, None, [7]),
    (25, None, [25]),
    (60, None, [25, 25, 10]),
    (35, 10, [10, 10, 5, 5, 5]),
    (4, 2, [2, 2]),
])
def test_epochs_are_run_in_spans_that_end_at_reports_and_resampling_points(monkeypatch, epochs, resample, spans):
    log = []
    plan = _Plan(log)
    model = mfm.MatrixFactorization(4, loss_graph=mfm.WMRBLoss(), n_users=5, n_items=7, n_samples=3)
    monkeypatch.setattr(model, "_prepare", lambda *a, **k: plan)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(mfm, "random_sampler", lambda *a, **k: np.zeros((5, 3), np.int64))
    model.fit(epochs, None, None, None, lr=0.1, verbose=False, resample_every=resample, resample_seed=1)
    runs = [n for kind, n in log if kind == "run"]
    assert runs == spans and sum(runs) == epochs
    # one loss report per completed 25 epochs, taken right after the span that ends there
    assert sum(1 for kind, _ in log if kind == "loss") == epochs // 25
    # steps done before each resampling are multiples of resample_every; the captured step is dropped right after it
    done = 0
    for j, (kind, n) in enumerate(log):
        if kind == "run":
            done += n
        if kind == "resample":
            assert resample and done % resample == 0 and 0 < done < epochs
            assert log[j + 1][0] == "invalidate"
    assert log[-1][0] == "invalidate"  # the captured step does not outlive fit()
    assert [e for e, _ in model.loss_history] == [25 * (q + 1) for q in range(epochs // 25)]
