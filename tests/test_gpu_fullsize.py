"""Parity at the FULL-SIZE networks of BASELINE.json configs 3, 4 and 5 -- needs a B200 (-m gpu).

The small differential tests (test_gpu_parity.py) never reach the code paths these networks take: the
grid.y split of a 10 000-node walk, the mixed population of dense (16 k-entry CPTs) and walked nodes of the
2 000-node DAG, resident-chunk sizing at 119 200 state values per case, and both 128-case accumulators of
the tcgen05 tile at K = 1024.  The oracle port runs a dozen cases of either network in seconds.

Tolerances (BASELINE.json north_star): fp64 |a-b| <= 1e-9*max(|a|,|b|) + 1e-12, fp32 1e-5 / 1e-7."""
import numpy as np
import pytest

from bayesiannetwork_b200 import synth
from helpers import assert_close

pytestmark = pytest.mark.gpu
TOL = {"fp64": dict(rtol=1e-9, atol=1e-12), "fp32": dict(rtol=1e-5, atol=1e-7)}


@pytest.fixture(scope="module")
def BP():
    from bayesiannetwork_b200.engine import BeliefPropagation
    return BeliefPropagation


_cache = {}


def _port(oracle_mod, key, net, ev, eps, cap):
    """One oracle run per (network, evidence, stopping rule), shared by the precision parametrisations."""
    if key not in _cache:
        _cache[key] = oracle_mod.run_port(net, ev, eps=eps, max_sweeps=cap, threads=0)
    return _cache[key]


@pytest.fixture(scope="module")
def grid100():
    return synth.grid(100)


@pytest.fixture(scope="module")
def dag2000():
    return synth.random_dag()


# ---- cfg 3: 100 x 100 binary grid, fixed 50 sweeps ------------------------------------------------------
@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_grid100_full_size_50_sweeps(BP, oracle_mod, grid100, precision):
    net = grid100
    assert net.n_nodes == 10000 and net.n_edges == 19800 and net.state_values == 119200
    ev = synth.make_evidence(net, 12, p=0.10, seed=5)              # 12 cases: one ragged tile, the node walk splits over grid.y
    om, osw, _ = _port(oracle_mod, "grid100", net, ev, 0.0, 50)
    res = BP(net, precision)(ev, 0.0, max_sweeps=50)
    assert np.array_equal(res.sweeps, osw) and res.sweeps[0] == 50
    assert_close(res.marginals, om, what=f"grid100 {precision}", **TOL[precision])


def test_grid100_resident_chunks_and_wide_batch(BP, oracle_mod, grid100):
    """More cases than stay resident (max_resident_cases): chunked and unchunked runs agree bit for bit, and a
    sample of the wide batch agrees with the oracle (the batch-wide launch shape: no grid.y split)."""
    net = grid100
    ev = synth.make_evidence(net, 1100, p=0.10, seed=6)
    a = BP(net, "fp64")(ev, 0.0, max_sweeps=6)
    b = BP(net, "fp64", max_resident_cases=512)(ev, 0.0, max_sweeps=6)
    assert np.array_equal(a.marginals, b.marginals)
    idx = [0, 1, 511, 512, 513, 1023, 1024, 1099]
    for i in idx:
        one = ev.slice(i, i + 1)
        om, _, _ = oracle_mod.run_port(net, one, eps=0.0, max_sweeps=6, threads=0)
        assert_close(a.marginals[i:i + 1], om, what=f"grid100 wide batch case {i}", **TOL["fp64"])


# ---- cfg 4: random DAG, 2 000 nodes, <= 4 parents, card 2-8 -------------------------------------------
@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_dag2000_full_size_20_sweeps(BP, oracle_mod, dag2000, precision):
    net = dag2000
    assert net.n_nodes == 2000 and int(net.card.max()) == 8 and int(np.diff(net.parent_off).max()) == 4
    ev = synth.make_evidence(net, 12, p=0.10, seed=7)
    om, osw, _ = _port(oracle_mod, "dag2000", net, ev, 0.0, 20)
    bp = BP(net, precision)
    st = bp.stats()
    assert 0 < st["dense_nodes"] < net.n_nodes                      # both node populations are present
    res = bp(ev, 0.0, max_sweeps=20)
    assert np.array_equal(res.sweeps, osw)
    assert_close(res.marginals, om, what=f"dag2000 {precision}", **TOL[precision])


def test_dag2000_full_size_epsilon_mode(BP, oracle_mod, dag2000):
    """The reference's own stopping rule at full size: equal sweep counts and flags, beliefs within 1e-9."""
    net = dag2000
    ev = synth.make_evidence(net, 10, p=0.10, seed=8)
    om, osw, ocv = _port(oracle_mod, "dag2000_eps", net, ev, 1e-6, 200)
    res = BP(net, "fp64")(ev, 1e-6, max_sweeps=200)
    assert np.array_equal(res.sweeps, osw), (res.sweeps, osw)
    assert np.array_equal(res.converged, ocv)
    assert_close(res.marginals, om, what="dag2000 eps", **TOL["fp64"])


def test_dag2000_walked_only_matches_dense(BP, dag2000):
    """dense_min_cpt = -1 walks every CPT in the sweep kernel: same beliefs as the default mixed path."""
    net = dag2000
    ev = synth.make_evidence(net, 140, p=0.10, seed=9)              # two tiles, the second ragged
    a = BP(net, "fp64")(ev, 0.0, max_sweeps=5)
    b = BP(net, "fp64", dense_min_cpt=-1)(ev, 0.0, max_sweeps=5)
    assert_close(a.marginals, b.marginals, what="dag2000 dense vs walked", **TOL["fp64"])


# ---- cfg 5: card 32, 3 parents: products of depth K = 1024 ----------------------------------------------
@pytest.fixture(scope="module")
def card32_k3():
    # the node shape of cfg 5 (32^4-entry CPTs, K = 1024) on 5 nodes, so that the oracle finishes 300 cases in
    # well under a minute; the 64-node network itself runs in test_gpu_dense_tc.py::test_card32_full_size_sample
    return synth.high_card(5, card=32, n_parents=3, seed=12)


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_card32_k1024_300_ragged_cases_10_sweeps(BP, oracle_mod, card32_k3, precision):
    """300 cases = two full 128-case accumulators in the first 256-case CTA and a ragged first / empty second
    one in the last; 10 sweeps."""
    net = card32_k3
    ev = synth.make_evidence(net, 300, p=0.15, seed=4)
    om, _, _ = _port(oracle_mod, "card32_k3", net, ev, 0.0, 10)
    bp = BP(net, precision)
    st = bp.stats()
    assert st["dense_nodes"] == 2
    if precision == "fp32":
        assert st["dense_tensor_jobs"] == 4                         # both products of both nodes on tcgen05
    assert_close(bp(ev, 0.0, max_sweeps=10).marginals, om, what=f"card32 K=1024 {precision}", **TOL[precision])


def test_card32_tensor_tiles_other_batch_shapes(BP, oracle_mod, card32_k3):
    """129 cases (second accumulator holds one row) and 257 (a second CTA with one row): tensor-core path vs the
    CUDA-core products of the same handle configuration."""
    net = card32_k3
    for n in (129, 257):
        ev = synth.make_evidence(net, n, p=0.15, seed=20 + n)
        a = BP(net, "fp32")(ev, 0.0, max_sweeps=4)
        b = BP(net, "fp32", dense_tensor=-1)(ev, 0.0, max_sweeps=4)
        assert_close(a.marginals, b.marginals, rtol=2e-5, atol=2e-7, what=f"card32 {n} cases tensor vs FMA")
        assert np.isfinite(a.marginals).all()


# ---- the asynchronous device path reports malformed evidence on request ---------------------------------
def test_device_path_error_flag(BP):
    import torch
    from bayesiannetwork_b200.engine import BnbpError
    net = synth.alarm37()
    bp = BP(net)
    dev = torch.device("cuda", 0)
    n, V = 256, net.belief_values
    off = torch.arange(n + 1, dtype=torch.int64, device=dev)
    node = torch.full((n,), 3, dtype=torch.int32, device=dev)
    state = torch.zeros(n, dtype=torch.int32, device=dev)
    out = torch.empty((n, V), dtype=torch.float64, device=dev)
    bad = node.clone()
    bad[17] = 99                                                   # node id out of range
    bp.run_device(n, off, bad, state, out, epsilon=0.0, max_sweeps=3)
    torch.cuda.synchronize()
    with pytest.raises(BnbpError):
        bp.check_errors()
    bp.check_errors()                                              # reported once, then clear
    # a flag nobody asked about does not fail the next call, on either path
    bp.run_device(n, off, bad, state, out, epsilon=0.0, max_sweeps=3)
    torch.cuda.synchronize()
    bp.run_device(n, off, node, state, out, epsilon=0.0, max_sweeps=3)
    torch.cuda.synchronize()
    bp.check_errors()
    bp.run_device(n, off, bad, state, out, epsilon=0.0, max_sweeps=3)
    torch.cuda.synchronize()
    res = bp(synth.make_evidence(net, 64, exact_k=4), 0.0, max_sweeps=3)
    assert np.isfinite(res.marginals).all()
    # epsilon mode synchronises anyway: the call itself fails
    with pytest.raises(BnbpError):
        bp.run_device(n, off, bad, state, out, epsilon=1e-6, max_sweeps=50)
