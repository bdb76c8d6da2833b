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
  * PINNED against the reference's OWN SOURCE executed here: the unmodified
    package ``/root/reference/src/teamoflow/mf`` is imported over a stand-in for
    the TensorFlow / TFP entry points it calls (``tests/golden/tf_shim``,
    torch-CPU fp32) and its outputs are committed as
    ``tests/golden/ref_golden.json`` (generator: ``tests/golden/make_ref_golden.py``):
    the three loss graphs, ``fit`` for 1 and 2 epochs over 9 loss / tower /
    feature combinations, the embedding graphs and the whole evaluation surface.
    ``tests/test_ref_golden.py`` holds the oracle to those vectors (rankings,
    metrics and ``predict`` bit for bit).
  * ``gather_matrix_indices``: also pinned against the reference's only
    known-answer vector (``test/test_utils.py:47-61``).
  * STILL UNPINNED: the TensorFlow kernels themselves.  The op semantics tagged
    [TF-sem] (``top_k`` tie order, ``maximum``'s sub-gradient at equality, Keras
    Adam's update rule, ``l2_normalize``'s epsilon, TFP's ``ndtr``) are restated
    from documentation in both the shim and this oracle; no TensorFlow run was
    possible.  Further cross-checks: (1) a pure-Python scalar-loop restatement
    written independently of the vectorised NumPy code
    (``tests/golden/make_golden.py``), (2) ``torch.autograd`` on CPU over the
    dense graph written op-for-op like the reference
    (``oracle/autograd_twin.py``), (3) hand-derived vectors for those
    semantics.
"""
from . import mf_oracle  # noqa: F401
