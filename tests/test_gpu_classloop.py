"""Parity of the CLASS-LOOPED specialised sweep kernels (bnbp_spec.cuh with BNBP_CLASSLOOP: one unrolled body per
node shape class, looped over the class's nodes) -- needs a B200 (-m gpu).

Three angles: (1) the small differential networks forced through the class-looped generator (BNBP_CLASSLOOP=2)
against the oracle, same bars as test_gpu_spec.py; (2) class-looped == unrolled bit for bit on the headline network
(same arithmetic per node, only the walk order and where the offsets come from differ); (3) the network the mode
exists for -- cfg 3, the 10 000-node grid -- at full size against the oracle and against the generic kernel.
The host-side emulation of the same generated source is tests/test_netcompiler_emul.py (no GPU)."""
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


def _hub():
    card = [3] + [2 + (i % 3) for i in range(12)] + [2, 4]
    parents = [[]] + [[0] for _ in range(12)] + [[], [5, 9]]
    return synth._assemble(card, parents, 77, "hub12")


def _cases():
    yield "polytree40_eps", synth.random_polytree(40, card_hi=4, seed=21), dict(p=0.15), 1e-8, 300
    yield "grid6_loopy_eps", synth.grid(6, seed=4), dict(p=0.1), 1e-7, 400
    yield "grid6_fixed", synth.grid(6, seed=4), dict(p=0.1), 0.0, 9
    yield "dag45_k4_fixed", synth.random_dag(45, 4, 2, 3, seed=7), dict(p=0.1), 0.0, 12
    yield "hub12_isolated_eps", _hub(), dict(p=0.3), 1e-9, 200
    yield "grid6_soft", synth.grid(6, seed=4), dict(p=0.1, soft=True), 0.0, 7


@pytest.mark.parametrize("name,net,evkw,eps,cap", list(_cases()), ids=[c[0] for c in _cases()])
@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_classloop_vs_oracle(BP, oracle_mod, monkeypatch, name, net, evkw, eps, cap, precision):
    if precision == "fp32" and eps > 0:
        pytest.skip("fp32 parity is asserted at fixed sweep counts")
    monkeypatch.setenv("BNBP_CLASSLOOP", "2")
    ev = synth.make_evidence(net, 777, seed=17, **evkw)          # ragged: not a multiple of any tile
    om, osw, ocv = oracle_mod.run_port(net, ev, eps=eps, max_sweeps=cap, threads=0)
    bp = BP(net, precision, specialize="always")
    res = bp(ev, eps, max_sweeps=cap)
    st = bp.stats()
    assert st["last_specialised"] == 1 and st["spec_class_count"] > 0 and st["last_fused"] == 0
    assert np.array_equal(res.sweeps, osw), (name, np.nonzero(res.sweeps != osw)[0][:5])
    assert np.array_equal(res.converged, ocv), name
    assert_close(res.marginals, om, what=name, **TOL[precision])


def test_classloop_equals_unrolled_bit_for_bit(BP, monkeypatch):
    """alarm37 (36 classes for 37 nodes: nothing to gain, everything to compare): a fixed-count run and an epsilon run
    through both code generators give the same bits -- the node arithmetic is the same template code."""
    net = synth.alarm37()
    ev = synth.make_evidence(net, 4096 + 33, exact_k=4, seed=23)
    monkeypatch.setenv("BNBP_ONCHIP", "0")
    res = {}
    for mode in ("0", "2"):
        monkeypatch.setenv("BNBP_CLASSLOOP", mode)
        bp = BP(net, "fp64", specialize="always")
        a = bp(ev, 0.0, max_sweeps=8)
        assert (bp.stats()["spec_class_count"] > 0) == (mode == "2") and bp.stats()["last_onchip"] == 0
        b = bp(ev, 1e-6, max_sweeps=200)
        res[mode] = (a, b)
    for i in (0, 1):
        assert np.array_equal(res["0"][i].marginals, res["2"][i].marginals)
        assert np.array_equal(res["0"][i].sweeps, res["2"][i].sweeps)
        assert np.array_equal(res["0"][i].converged, res["2"][i].converged)


# ---- cfg 3: the 100 x 100 grid ---------------------------------------------------------------------------
@pytest.fixture(scope="module")
def grid100():
    return synth.grid(100)


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_grid100_class_looped_50_sweeps(BP, oracle_mod, grid100, precision):
    net = grid100
    ev = synth.make_evidence(net, 12, p=0.10, seed=5)
    om, osw, _ = oracle_mod.run_port(net, ev, eps=0.0, max_sweeps=50, threads=0)
    bp = BP(net, precision, specialize="always")                  # "always": a 12-case batch would take the generic kernel
    res = bp(ev, 0.0, max_sweeps=50)
    st = bp.stats()
    assert st["last_specialised"] == 1 and 4 <= st["spec_class_count"] <= 9 and st["cases_per_tile"] in (128, 256)
    assert np.array_equal(res.sweeps, osw) and res.sweeps[0] == 50
    assert_close(res.marginals, om, what=f"grid100 class-looped {precision}", **TOL[precision])


def test_grid100_class_looped_epsilon_mode(BP, oracle_mod, grid100):
    net = grid100
    ev = synth.make_evidence(net, 10, p=0.10, seed=8)
    om, osw, ocv = oracle_mod.run_port(net, ev, eps=1e-6, max_sweeps=400, threads=0)
    bp = BP(net, "fp64", specialize="always")
    res = bp(ev, 1e-6, max_sweeps=400)
    assert bp.stats()["spec_class_count"] > 0
    assert np.array_equal(res.sweeps, osw) and np.array_equal(res.converged, ocv)
    assert_close(res.marginals, om, what="grid100 class-looped eps", **TOL["fp64"])


def test_grid100_wide_batch_default_path_is_class_looped(BP, oracle_mod, grid100):
    """What bench.py's cfg 3 entry runs: >= 4096 cases under the default options take the class-looped kernel; the
    whole batch agrees with the generic kernel to rounding, a sample agrees with the oracle, and a chunked run (cases
    that do not stay resident) agrees with the unchunked one bit for bit."""
    net = grid100
    ev = synth.make_evidence(net, 4096 + 77, p=0.10, seed=6)
    bp = BP(net, "fp64")
    a = bp(ev, 0.0, max_sweeps=6)
    st = bp.stats()
    assert st["last_specialised"] == 1 and st["spec_class_count"] > 0
    g = BP(net, "fp64", specialize="never")(ev, 0.0, max_sweeps=6)
    assert_close(a.marginals, g.marginals, what="grid100 class-looped vs generic", **TOL["fp64"])
    b = BP(net, "fp64", max_resident_cases=1536)(ev, 0.0, max_sweeps=6)
    assert np.array_equal(a.marginals, b.marginals)
    for i in (0, 127, 128, 4095, 4096, 4172):
        om, _, _ = oracle_mod.run_port(net, ev.slice(i, i + 1), eps=0.0, max_sweeps=6, threads=0)
        assert_close(a.marginals[i:i + 1], om, what=f"grid100 wide batch case {i}", **TOL["fp64"])


def test_grid100_small_batch_takes_node_slices(BP, oracle_mod, monkeypatch, grid100):
    """cfg 3 sharded over several GPUs leaves a few thousand cases per device: fewer tiles than SMs.  The plain variants
    then run with node slices in grid.y (block (tile, y) walks the y-th slice of every class, one sweep per launch);
    sliced == unsliced bit for bit, and both agree with the oracle."""
    net = grid100
    ev = synth.make_evidence(net, 1100, p=0.10, seed=6)           # 9 tiles -> 32 slices
    bp = BP(net, "fp64", specialize="always")
    a = bp(ev, 0.0, max_sweeps=7)
    st = bp.stats()
    assert st["spec_class_count"] > 0 and st["last_sweep_launches"] >= 7        # one sweep per launch (per chunk of the call)
    monkeypatch.setenv("BNBP_NO_NODE_SLICES", "1")
    b = bp(ev, 0.0, max_sweeps=7)
    assert np.array_equal(a.marginals, b.marginals)
    monkeypatch.delenv("BNBP_NO_NODE_SLICES")
    monkeypatch.setenv("BNBP_NODE_SLICES", "5")
    c = bp(ev, 0.0, max_sweeps=7)
    assert np.array_equal(a.marginals, c.marginals)
    for i in (0, 511, 1099):
        om, _, _ = oracle_mod.run_port(net, ev.slice(i, i + 1), eps=0.0, max_sweeps=7, threads=0)
        assert_close(a.marginals[i:i + 1], om, what=f"grid100 sliced case {i}", **TOL["fp64"])
