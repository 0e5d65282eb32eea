"""The dense contraction path (csrc/bnbp_dense.cuh: CPT x batch as two matrix products per node and
sweep, finished by the sweep kernel from per-case tables) against the oracle and the reference
fixtures -- needs a B200 (-m gpu).

`dense_min_cpt` is lowered so that small networks exercise every split (s = 1..k, group B empty or
not), both tile widths (128 / 256 cases per state tile), ragged batches, evidence on dense nodes,
NaN propagation, soft evidence, epsilon mode with frozen tiles, damping and CPT refresh."""
import numpy as np
import pytest

from bayesiannetwork_b200 import synth
from bayesiannetwork_b200.flat import EvidenceBatch, FlatNetwork
from helpers import assert_close, load_fixture
from test_gpu_parity import FIXTURES, TOL

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def BP():
    from bayesiannetwork_b200.engine import BeliefPropagation
    return BeliefPropagation


@pytest.mark.parametrize("name", FIXTURES)
def test_reference_fixtures_all_nodes_dense(BP, ref_fixtures, name):
    """Every non-root node takes the dense path (threshold 2 entries); outputs of the unmodified
    reference are the bar, including sweep counts and the NaN pattern of impossible evidence."""
    f = load_fixture(ref_fixtures, name)
    bp = BP(f["net"], "fp64", dense_min_cpt=2)
    assert bp.stats()["dense_nodes"] > 0
    res = bp(f["ev"], f["eps"], max_sweeps=f["max_sweeps"])
    assert bp.stats()["last_dense_launches"] > 0
    assert np.array_equal(res.sweeps, f["sweeps"]), (name, res.sweeps, f["sweeps"])
    assert np.array_equal(res.converged, f["converged"]), name
    assert_close(res.marginals, f["marginals"], what=name, **TOL["fp64"])


def _cases():
    # (name, network, evidence kwargs, eps, sweep cap, dense_min_cpt)
    yield "card8_k3_default_threshold", synth.high_card(12, card=8, n_parents=3, seed=31), dict(p=0.15), 0.0, 6, 0
    yield "card8_k3_eps", synth.high_card(12, card=8, n_parents=3, seed=31), dict(p=0.15), 1e-7, 200, 0
    yield "card16_k2", synth.high_card(10, card=16, n_parents=2, seed=9), dict(p=0.2), 0.0, 6, 256
    yield "card32_k2", synth.high_card(6, card=32, n_parents=2, seed=10), dict(p=0.2), 0.0, 5, 1024
    yield "card12_k3", synth.random_dag(24, 3, 2, 12, seed=8), dict(p=0.1), 0.0, 8, 64
    yield "dag120_card5_all", synth.random_dag(120, 4, 2, 5, seed=6), dict(p=0.1), 0.0, 12, 2
    yield "dag80_card8_mixed_eps", synth.random_dag(80, 4, 2, 8, seed=7), dict(p=0.1), 1e-7, 300, 200
    yield "dag80_card8_all", synth.random_dag(80, 4, 2, 8, seed=7), dict(p=0.1), 0.0, 9, 2
    yield "wide_parents_k7", synth.random_dag(40, 7, 2, 3, seed=12), dict(p=0.1), 0.0, 6, 16
    yield "alarm37_vec2_eps", synth.alarm37(), dict(exact_k=4), 1e-6, 200, 2
    yield "alarm37_soft", synth.alarm37(), dict(exact_k=4, soft=True), 0.0, 15, 8
    yield "grid8", synth.grid(8, seed=4), dict(p=0.1), 0.0, 30, 2


@pytest.mark.parametrize("name,net,evkw,eps,cap,dmin", list(_cases()), ids=[c[0] for c in _cases()])
@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_dense_vs_oracle(BP, oracle_mod, name, net, evkw, eps, cap, dmin, precision):
    if precision == "fp32" and eps > 0:
        pytest.skip("fp32 parity is asserted at fixed sweep counts")
    ev = synth.make_evidence(net, 300, seed=17, **evkw)
    om, osw, ocv = oracle_mod.run_port(net, ev, eps=eps, max_sweeps=cap, threads=0)
    bp = BP(net, precision, dense_min_cpt=dmin)
    assert bp.stats()["dense_nodes"] > 0, name
    res = bp(ev, eps, max_sweeps=cap)
    assert np.array_equal(res.sweeps, osw), (name, np.nonzero(res.sweeps != osw)[0][:5])
    assert np.array_equal(res.converged, ocv), name
    assert_close(res.marginals, om, what=name, **TOL[precision])


@pytest.mark.parametrize("n_cases", [1, 127, 129, 1000])
def test_dense_ragged_batches_and_resident_chunks(BP, oracle_mod, n_cases):
    net = synth.random_dag(50, 4, 2, 6, seed=41)
    ev = synth.make_evidence(net, n_cases, p=0.15, seed=5)
    om, osw, _ = oracle_mod.run_port(net, ev, eps=1e-8, max_sweeps=150, threads=0)
    res = BP(net, dense_min_cpt=2)(ev, 1e-8, max_sweeps=150)
    assert np.array_equal(res.sweeps, osw)
    assert_close(res.marginals, om, what=f"ragged{n_cases}", **TOL["fp64"])
    if n_cases == 1000:      # the same batch cut into resident chunks of 512 cases
        res2 = BP(net, dense_min_cpt=2, max_resident_cases=512)(ev, 1e-8, max_sweeps=150)
        assert np.array_equal(res2.sweeps, res.sweeps) and np.array_equal(res2.marginals, res.marginals)


def test_dense_equals_per_thread_path(BP):
    """Same network and evidence through both formulations of the parent side (4|CPT| flops via the
    matrix products vs the one-pass recursion): equal up to reassociation."""
    net = synth.high_card(16, card=8, n_parents=3, seed=3)
    ev = synth.make_evidence(net, 2048, p=0.1, seed=9)
    a = BP(net, dense_min_cpt=-1)
    b = BP(net)
    assert a.stats()["dense_nodes"] == 0 and b.stats()["dense_nodes"] == 13
    ra, rb = a(ev, 0.0, max_sweeps=10), b(ev, 0.0, max_sweeps=10)
    assert_close(ra.marginals, rb.marginals, rtol=1e-11, atol=1e-14, what="dense vs per-thread")
    st = b.stats()
    # both products of a card-8 / 3-parent node have two factors: one launch per sweep
    assert st["last_dense_launches"] == 10 and st["dense_values_per_case"] == 13 * 2 * 64
    assert st["dense_flops_per_case_sweep"] == 4.0 * 13 * 8 ** 4


def test_dense_extensions_and_refresh(BP, oracle_mod):
    net = synth.random_dag(60, 4, 2, 6, seed=13)
    ev = synth.make_evidence(net, 400, p=0.1, seed=2)
    bp = BP(net, dense_min_cpt=2)
    om, osw, ocv = oracle_mod.run_port(net, ev, eps=1e-7, max_sweeps=300, damping=0.3, check_interval=3, threads=0)
    res = bp(ev, 1e-7, max_sweeps=300, damping=0.3, check_interval=3)
    assert np.array_equal(res.sweeps, osw) and np.array_equal(res.converged, ocv)
    assert_close(res.marginals, om, what="damped", **TOL["fp64"])
    # new CPT values, same topology: both arenas (reference layout and the transposed copies) refresh
    rng = np.random.default_rng(5)
    cpt = net.cpt.copy()
    for x in range(net.n_nodes):
        rows = cpt[net.cpt_off[x]:net.cpt_off[x + 1]].reshape(-1, int(net.card[x]))
        rows[:] = 0.05 + rng.random(rows.shape)
        rows /= rows.sum(axis=1, keepdims=True)
    net2 = FlatNetwork(net.card, net.parent_off, net.parents, net.cpt_off, cpt, name="refreshed")
    bp.refresh_cpt(net2.cpt)
    om2, _, _ = oracle_mod.run_port(net2, ev, eps=0.0, max_sweeps=10, threads=0)
    assert_close(bp(ev, 0.0, max_sweeps=10).marginals, om2, what="refreshed", **TOL["fp64"])
