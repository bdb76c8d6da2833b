"""Host logic of ``TrainPlan.run`` (first step eager, the rest replayed from one captured graph; eager fallbacks; the
launch accounting bench.py's ``gpu_launches`` claim relies on).  CPU only: ``torch.cuda.CUDAGraph`` / ``torch.cuda.graph``
are replaced by recording stand-ins."""
import pytest
import torch

from teamoflow_b200 import _abi
from teamoflow_b200.mf import _engine as eng

LAUNCHES_PER_STEP = 14


class _Graph:
    replays = 0

    def replay(self):
        _Graph.replays += 1


def _plan(monkeypatch, capture_fails=False):
    plan = object.__new__(eng.TrainPlan)
    plan.opt_state = None
    plan.comm = None
    plan.batched = False
    calls = {"eager": 0, "captured": 0, "capturing": False}

    def step(lr):
        calls["captured" if calls["capturing"] else "eager"] += 1
        _abi.launch_count += LAUNCHES_PER_STEP
        _abi.call_count += LAUNCHES_PER_STEP

    def capture(fn):  # stand-in for eng.capture_graph (stream + capture_begin / capture_end around fn)
        if capture_fails:
            raise RuntimeError("capture refused")
        calls["capturing"] = True
        try:
            fn()
        finally:
            calls["capturing"] = False
        return _Graph()

    plan.step = step
    _Graph.replays = 0
    monkeypatch.setattr(eng, "capture_graph", capture)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(eng.TrainPlan, "USE_CUDA_GRAPH", True)
    return plan, calls


def test_first_step_eager_then_replays_and_launches_are_counted_once_per_step(monkeypatch):
    plan, calls = _plan(monkeypatch)
    l0 = _abi.launch_count
    plan.run(5, 0.1)
    assert calls == {"eager": 1, "captured": 1, "capturing": False} and _Graph.replays == 4
    assert _abi.launch_count - l0 == 5 * LAUNCHES_PER_STEP  # the capture itself launches nothing
    plan.run(3, 0.1)  # same lr: the captured step is reused
    assert calls["captured"] == 1 and _Graph.replays == 6 and plan.graph_replays == 6
    plan.run(3, 0.05)  # another lr is another graph (lr is baked into the update launches)
    assert calls["captured"] == 2
    plan.invalidate_graph()
    plan.run(3, 0.05)
    assert calls["captured"] == 3


@pytest.mark.parametrize("case", ["short", "stateful_adam", "switched_off", "argument"])
def test_eager_paths(monkeypatch, case):
    plan, calls = _plan(monkeypatch)
    n = 2 if case == "short" else 6
    if case == "stateful_adam":
        plan.opt_state = ({}, {})  # the step count is a per-call argument: not replayable
    if case == "switched_off":
        monkeypatch.setattr(eng.TrainPlan, "USE_CUDA_GRAPH", False)
    plan.run(n, 0.1, graph=False if case == "argument" else None)
    assert calls["eager"] == n and calls["captured"] == 0 and _Graph.replays == 0


def test_refused_capture_falls_back_to_the_eager_loop(monkeypatch, capsys):
    plan, calls = _plan(monkeypatch, capture_fails=True)
    l0 = _abi.launch_count
    plan.run(4, 0.1)
    assert calls["eager"] == 4 and _Graph.replays == 0 and _abi.launch_count - l0 == 4 * LAUNCHES_PER_STEP
    assert "capture of the training step failed" in capsys.readouterr().err
    # the failure is recorded on THIS plan only: other plans of the process may still capture
    assert plan._graph_failed is True and eng.TrainPlan.USE_CUDA_GRAPH is True
    plan.run(4, 0.1)  # not retried
    assert calls["eager"] == 8 and _Graph.replays == 0
