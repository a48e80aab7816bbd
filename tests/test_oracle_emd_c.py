"""The C restatement of the exact transport LP (oracle/emd_netsimplex.c: network simplex, the algorithm class of POT's `ot.emd2`,
FilteringMergingModule.py:160-166) against the oracle's other exact solvers.  Test infrastructure checking test infrastructure:
the product never calls any of them."""
import numpy as np
import pytest
import torch

from oracle import mars_oracle as orc
from oracle.emd_c import emd_network_simplex_c


@pytest.mark.parametrize("t,m", [(1, 1), (1, 9), (9, 1), (2, 3), (7, 7), (16, 16), (13, 40), (40, 13), (35, 36), (64, 27), (90, 55)])
def test_network_simplex_c_equals_highs(t, m):
    rng = np.random.default_rng(100 * t + m)
    c = rng.random((t, m)).astype(np.float32).astype(np.float64)
    got, got_int, pivots = emd_network_simplex_c(c, details=True)
    assert abs(got - orc.emd_exact(c)) < 1e-12
    assert abs(got - got_int) < 1e-12  # float32 costs above 2^-17 are exact on the 2^-40 grid
    assert pivots >= 0


@pytest.mark.parametrize("case", ["all_equal", "zeros", "three_values", "assignment", "block_structure", "duplicate_columns"])
def test_network_simplex_c_on_degenerate_problems(case):
    """Ties everywhere: the perturbed marginals keep every basis non-degenerate, so the method terminates, and the optimum of the
    original problem comes out exact."""
    rng = np.random.default_rng(7)
    if case == "all_equal":
        c = np.full((12, 18), 0.375)
        assert abs(emd_network_simplex_c(c) - 0.375) < 1e-15
        return
    if case == "zeros":
        assert emd_network_simplex_c(np.zeros((9, 6))) == 0.0
        return
    if case == "three_values":
        c = rng.choice([0.0, 0.5, 1.0], size=(24, 30))
    elif case == "assignment":  # T = M: the LP optimum is an assignment, every basic solution of the original problem is degenerate
        from scipy.optimize import linear_sum_assignment

        c = rng.random((30, 30)).astype(np.float32).astype(np.float64)
        r, k = linear_sum_assignment(c)
        assert abs(emd_network_simplex_c(c) - c[r, k].sum() / 30) < 1e-12
    elif case == "block_structure":  # the synthetic features' prototypes give block-structured costs
        a, b = rng.integers(0, 4, 40), rng.integers(0, 4, 28)
        c = np.where(a[:, None] == b[None, :], 0.1, 0.6) + 0.01 * rng.random((40, 28))
        c = c.astype(np.float32).astype(np.float64)
    else:
        base = rng.random((20, 6)).astype(np.float32).astype(np.float64)
        c = base[:, rng.integers(0, 6, 25)]
    assert abs(emd_network_simplex_c(c) - orc.emd_exact(c)) < 1e-12


def test_network_simplex_c_equals_networkx_and_the_expanded_assignment():
    """Two more independent exact solvers: networkx's network simplex on the float32 grid and the lcm(T, M)^2 assignment."""
    from math import lcm

    from scipy.optimize import linear_sum_assignment

    rng = np.random.default_rng(3)
    for t, m in ((6, 4), (9, 12), (10, 15)):
        c = (2.0 ** -5 + (1 - 2.0 ** -5) * rng.random((t, m))).astype(np.float32).astype(np.float64)
        got = emd_network_simplex_c(c)
        assert abs(got - orc.emd_network_simplex(c)) < 1e-13
        n = lcm(t, m)
        big = np.repeat(np.repeat(c, n // t, axis=0), n // m, axis=1)
        r, k = linear_sum_assignment(big)
        assert abs(got - big[r, k].sum() / n) < 1e-12


def test_network_simplex_c_on_the_lps_of_a_synthetic_episode():
    """The LPs the path actually poses: foreground support rows against the pooled patches of every proposal of an episode
    (cosine costs with prototype structure, duplicated and near-duplicated proposals, T and M mostly coprime)."""
    import marsb200

    shape = marsb200.EpisodeShape(ns=1, g=12, C=64, P=12, H=96, W=96, gt=8, D=16)
    ep = marsb200.make_episode(shape, 31)
    fs, fq = orc.normalize_rows(ep["feat_s"].reshape(-1, shape.C)), orc.normalize_rows(ep["feat_q"])
    _, cost = orc.similarity_and_cost(fs, fq)
    sup = orc.pool_mask(ep["support_mask"].float(), shape.g).reshape(-1)
    pm = orc.pool_mask(ep["masks"].float(), shape.g).reshape(shape.P, -1)
    for p in range(shape.P):
        sub = cost[sup.bool()][:, pm[p]].numpy().astype(np.float64)
        want = 1.0 - orc.emd_score(sup, pm[p], cost)
        assert abs(emd_network_simplex_c(sub) - want) < 1e-9


def test_network_simplex_c_edge_cases():
    assert emd_network_simplex_c(np.zeros((0, 5))) == 0.0  # an empty marginal is defined as zero cost
    assert emd_network_simplex_c(np.zeros((4, 0))) == 0.0
    with pytest.raises(RuntimeError):
        emd_network_simplex_c(np.array([[0.1, -0.2]]))
    with pytest.raises(RuntimeError):
        emd_network_simplex_c(np.array([[0.1, np.nan]]))
    with pytest.raises(ValueError):
        emd_network_simplex_c(np.zeros(5))
    c = torch.rand(5, 7).numpy()[:, ::2]  # non-contiguous float32 view
    assert abs(emd_network_simplex_c(c) - orc.emd_exact(np.ascontiguousarray(c, dtype=np.float64))) < 1e-12
