"""CPU oracle for the TeAMOFlow matrix-factorization hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``teamoflow_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and there only as the checker
or the timed CPU baseline -- never as the thing shipped.

What it is: a NumPy restatement of the reference's algorithm
(``/root/reference/src/teamoflow/mf``), each function citing the reference
file:line it follows.  The arithmetic of the reference lives in third-party,
un-vendored, *unpinned* dependencies (TensorFlow >= 2.9, tensorflow-probability
>= 0.17 -- ``README.md:19-27``; ``pyproject.toml`` declares no dependencies and
there is no lock file).  Neither is installed in this image and neither can be
installed (no network), so the reference itself cannot be executed here.

PARITY PINNING STATUS
  * ``gather_matrix_indices``: PINNED against the reference's own known-answer
    vector (``test/test_utils.py:47-61``), see ``tests/golden/``.
  * everything else (losses, gradients, Adam step, scores, top-k order,
    recall/precision/f1/ndcg, initializer statistics): **parity unpinned** --
    the reference's tests hold no golden values for them (they are
    exception-swallowing smoke tests) and TensorFlow cannot be run here.  The
    oracle is instead cross-checked three ways: (1) a pure-Python scalar-loop
    restatement written independently of the vectorised NumPy code
    (``tests/golden/make_golden.py``), (2) ``torch.autograd`` on CPU over the
    dense graph written op-for-op like the reference
    (``oracle/autograd_twin.py``), (3) hand-derived vectors for the documented
    TensorFlow semantics (sub-gradient of ``maximum`` at 0, ``top_k`` tie order,
    Adam step 1).
"""
from . import mf_oracle  # noqa: F401
