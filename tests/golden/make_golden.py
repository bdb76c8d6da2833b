"""Generates ``tests/golden/golden.json``.

Run from the repo root:  ``python tests/golden/make_golden.py``

The reference (TensorFlow) cannot be imported in this image, so these vectors are NOT
outputs of the reference.  They come from three independent sources:
  (a) the one known-answer vector the reference's own tests hold
      (``/root/reference/test/test_utils.py:47-61``), copied as literals;
  (b) a pure-Python scalar-loop restatement of the reference formulas written here,
      independently of the vectorised NumPy oracle (``oracle/mf_oracle.py``) -- fp64;
  (c) exact integer arithmetic for the top-k tie-order sets.
Everything in (b)/(c) is "parity unpinned" with respect to TensorFlow itself.
"""
from __future__ import annotations

import json
import math
import os
import random

HERE = os.path.dirname(os.path.abspath(__file__))


def dot(a, b):
    s = 0.0
    for x, y in zip(a, b):
        s += x * y
    return s


def phi(z):
    return math.exp(-0.5 * z * z) / math.sqrt(2 * math.pi)


def ndtr(z):
    return 0.5 * math.erfc(-z / math.sqrt(2.0))


# ---------------------------------------------------------------- scalar-loop model step


def step_scalar(loss, U, V, inter, samp=None, n_items=None, n_samples=None):
    """Identity features, LinearEmbedding: E_u = U, E_i = V.  Returns loss list, dU, dV (fp64).

    Follows loss_graphs.py:47-52 / :74-88 / :111-122 and the gradient of the SUM of the
    loss vector (matrix_factorization.py:170-171).
    """
    n_u, n_i, r = len(U), len(V), len(U[0])
    dU = [[0.0] * r for _ in range(n_u)]
    dV = [[0.0] * r for _ in range(n_i)]
    P = [[dot(U[u], V[i]) for i in range(n_i)] for u in range(n_u)]
    dP = [[0.0] * n_i for _ in range(n_u)]
    losses = []
    if loss == "mse":
        for (u, i, a) in inter:
            losses.append((a - P[u][i]) ** 2)
            dP[u][i] += -2.0 * (a - P[u][i])
    elif loss == "wmrb":
        scale = n_items / n_samples
        for (u, i, a) in inter:
            if not a > 0:
                continue
            hs = [1.0 - P[u][i] + P[u][j] for j in samp[u]]
            m = scale * sum(max(h, 0.0) for h in hs)
            losses.append(math.log(1.0 + m))
            w = scale / (1.0 + m)
            for j, h in zip(samp[u], hs):
                if h >= 0.0:  # TF: maximum(x, 0) sends the gradient to x when x >= 0
                    dP[u][j] += w
                    dP[u][i] -= w
    elif loss == "kl":
        pos = [(u, i) for (u, i, a) in inter if a > 0]
        neg = [(u, i) for (u, i, a) in inter if a <= 0]
        pp = [P[u][i] for u, i in pos]
        pn = [P[u][i] for u, i in neg]
        mp, mn = sum(pp) / len(pp), sum(pn) / len(pn)
        vp = sum((x - mp) ** 2 for x in pp) / len(pp)
        vn = sum((x - mn) ** 2 for x in pn) / len(pn)
        s = math.sqrt(vp + vn)
        z = (mp - mn) / s
        losses = 1.0 - ndtr(z)
        for (u, i), x in zip(pos, pp):
            dP[u][i] += -phi(z) / (s * len(pp)) * (1.0 - z * (x - mp) / s)
        for (u, i), x in zip(neg, pn):
            dP[u][i] += -phi(z) / (s * len(pn)) * (-1.0 - z * (x - mn) / s)
    for u in range(n_u):
        for i in range(n_i):
            g = dP[u][i]
            if g != 0.0:
                for c in range(r):
                    dU[u][c] += g * V[i][c]
                    dV[i][c] += g * U[u][c]
    return losses, dU, dV


def adam1(w, g, lr, b1=0.9, b2=0.999, eps=1e-7):
    m = (1 - b1) * g
    v = (1 - b2) * g * g
    alpha = lr * math.sqrt(1 - b2) / (1 - b1)
    return w - alpha * m / (math.sqrt(v) + eps)


# ---------------------------------------------------------------- exact top-k sets


def topk_exact(scores, k):
    """scores: list of python ints/Fractions; descending, ties -> lower index."""
    order = sorted(range(len(scores)), key=lambda i: (-scores[i], i))
    return order[:k]


def main():
    rnd = random.Random(1234)
    G = {}

    # (a) the reference's own known-answer vector, test/test_utils.py:47-61
    G["gather_matrix_indices"] = {
        "source": "reference test/test_utils.py:47-61",
        "input": [[1, 4, 2], [5, 7, 8], [6, 2, 1]],
        "index": [[0, 2, 0], [2, 2, 2], [2, 1, 0]],
        "expected": [[1, 2, 1], [8, 8, 8], [1, 2, 6]],
    }

    # (b1) WMRB 3 users x 4 items, S=2, r=2, dyadic values.
    #   user 0: sample list contains its own positive item (h = 1 exactly)
    #   user 1: an exact-zero hinge (1 - p + s == 0) -> tests the ">=" sub-gradient
    #   user 2: one positive, one NEGATIVE-valued interaction (ignored by WMRB)
    U = [[1.0, 0.5], [2.0, 0.0], [0.5, 0.5]]
    V = [[1.0, 1.0], [0.5, -1.0], [1.0, 0.0], [-0.5, 0.25]]
    #   p(1,0)=2, p(1,1)=1, p(1,2)=2, p(1,3)=-1 ;  positive (1,0): h vs sample 1 = 1-2+1 = 0 (exact)
    inter = [(0, 1, 3.0), (0, 2, 1.0), (1, 0, 5.0), (2, 0, -2.0), (2, 3, 4.0)]
    samp = [[1, 3], [1, 3], [0, 2]]
    losses, dU, dV = step_scalar("wmrb", U, V, inter, samp, n_items=4, n_samples=2)
    G["wmrb_3x4"] = {"U": U, "V": V, "inter": inter, "samp": samp, "n_items": 4, "n_samples": 2,
                     "loss": losses, "dU": dU, "dV": dV}

    # (b2) MSE 2x2
    U2 = [[0.5, -1.0, 0.25], [1.5, 0.5, -0.5]]
    V2 = [[1.0, 0.5, 2.0], [-0.5, 0.25, 1.0]]
    inter2 = [(0, 0, 1.0), (0, 1, -2.0), (1, 1, 3.0)]
    l2, dU2, dV2 = step_scalar("mse", U2, V2, inter2)
    G["mse_2x2"] = {"U": U2, "V": V2, "inter": inter2, "loss": l2, "dU": dU2, "dV": dV2}

    # (b3) KL, 2 positive / 2 non-positive
    U3 = [[0.5, 1.0], [1.0, -0.5], [0.25, 0.75]]
    V3 = [[1.0, 0.5], [-1.0, 0.5], [0.5, 0.5]]
    inter3 = [(0, 0, 2.0), (0, 1, -1.0), (1, 2, 4.0), (2, 1, -3.0)]
    l3, dU3, dV3 = step_scalar("kl", U3, V3, inter3)
    G["kl_2p2n"] = {"U": U3, "V": V3, "inter": inter3, "loss": l3, "dU": dU3, "dV": dV3}

    # (b4) random small cases (fp64 scalar loops) for each loss
    for name, loss in (("rand_mse", "mse"), ("rand_wmrb", "wmrb"), ("rand_kl", "kl")):
        n_u, n_i, r, S = 7, 9, 5, 4
        Ur = [[rnd.uniform(-1, 1) for _ in range(r)] for _ in range(n_u)]
        Vr = [[rnd.uniform(-1, 1) for _ in range(r)] for _ in range(n_i)]
        cells = [(u, i) for u in range(n_u) for i in range(n_i)]
        rnd.shuffle(cells)
        cells = sorted(cells[:25])
        it = [(u, i, float(rnd.choice([-3, -1, 1, 2, 5]))) for (u, i) in cells]
        sp = [rnd.sample(range(n_i), S) for _ in range(n_u)]
        lo, du, dv = step_scalar(loss, Ur, Vr, it, sp, n_items=n_i, n_samples=S)
        G[name] = {"U": Ur, "V": Vr, "inter": it, "samp": sp, "n_items": n_i, "n_samples": S,
                   "loss": lo, "dU": du, "dV": dv}

    # (b5) Adam step 1 table (fresh optimizer every step, matrix_factorization.py:176)
    gs = [0.0, 1e-8, -1e-8, 3.16e-6, -3.16e-6, 1.0, -1.0, 0.37]
    G["adam1"] = {"lr": 0.1, "w": 0.25, "g": gs, "expected": [adam1(0.25, g, 0.1) for g in gs]}

    # (b6) metrics: negative-valued interactions count as hits, all-non-positive score row,
    #      a user with no positives (preserve_rows -> x/0)
    P = [[0.9, 0.1, 0.5, -0.2, 0.3],   # top-2 = [0, 2]
         [-1.0, -0.5, -0.1, -2.0, -3.0],  # clamp -> all zeros -> top-2 = [0, 1]
         [0.2, 0.2, 0.7, 0.2, 0.1],   # tie at 0.2 -> top-2 = [2, 0]
         [0.4, 0.6, 0.1, 0.0, 0.3]]   # top-2 = [1, 0]
    A = [[1.0, 0.0, -2.0, 0.0, 3.0],    # relevant 2, hits {0:1, 2:-2} = 2 (negative counts)
         [0.0, 5.0, 0.0, 0.0, 0.0],     # relevant 1, hits {0:0, 1:5} = 1
         [0.0, 0.0, 0.0, 0.0, 0.0],     # relevant 0, hits 0  -> NaN -> 0 / dropped
         [-1.0, 0.0, 0.0, 0.0, 0.0]]    # relevant 0, hits {1:0, 0:-1} = 1 -> inf
    G["metrics_4x5"] = {
        "P": P, "A": A, "k": 2,
        "topk_clamped": [[0, 2], [0, 1], [2, 0], [1, 0]],
        "recall_drop": [1.0, 1.0],
        "recall_keep": [1.0, 1.0, 0.0, "inf"],
        "precision_drop": [1.0, 0.5],
        "precision_keep": [1.0, 0.5, 0.0, 0.5],
    }

    # (c) exact-grid top-k: embeddings on the k/64 grid => every partial sum is exactly
    #     representable in fp32 in any order; many ties; expected order from integer math.
    n_u, n_i, r, k = 24, 300, 16, 10
    Ug = [[rnd.randint(-8, 8) for _ in range(r)] for _ in range(n_u)]
    Vg = [[rnd.randint(-8, 8) for _ in range(r)] for _ in range(n_i)]
    for i in range(0, n_i, 7):  # duplicate item vectors -> exact ties
        Vg[i] = list(Vg[(i * 3) % n_i])
    raw, clamped = [], []
    for u in range(n_u):
        sc = [sum(a * b for a, b in zip(Ug[u], Vg[i])) for i in range(n_i)]
        raw.append(topk_exact(sc, k))
        clamped.append(topk_exact([s if s > 0 else 0 for s in sc], k))
    Ug[3] = [0] * r          # all-zero user row: every score ties at 0
    raw[3] = list(range(k))
    clamped[3] = list(range(k))
    Ug[5] = [-abs(x) for x in Ug[5]]
    sc5 = [sum(a * b for a, b in zip(Ug[5], Vg[i])) for i in range(n_i)]
    raw[5] = topk_exact(sc5, k)
    clamped[5] = topk_exact([s if s > 0 else 0 for s in sc5], k)
    G["grid_topk"] = {"scale": 64, "U_int": Ug, "V_int": Vg, "k": k, "raw": raw, "clamped": clamped}

    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(G, f, indent=0)
    print("wrote", os.path.join(HERE, "golden.json"))


if __name__ == "__main__":
    main()
