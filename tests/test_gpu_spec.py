"""Parity of the NETWORK-SPECIALISED sweep kernels (bnbp_spec.cuh, compiled by NVRTC on the box)
with the oracle and with the generic kernel -- needs a B200 (-m gpu).

Same bar as test_gpu_parity.py: fp64 1e-9 / 1e-12 with equal sweep counts in eps mode, fp32
1e-5 / 1e-7 at fixed sweep counts.  ``specialize="always"`` makes a missing / failing network
compiler an error instead of a silent switch to the generic kernel."""
import numpy as np
import pytest

from bayesiannetwork_b200 import synth
from bayesiannetwork_b200.flat import EvidenceBatch
from helpers import assert_close, load_fixture

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
    yield "alarm37_eps", synth.alarm37(), dict(exact_k=4), 1e-6, 200
    yield "alarm37_fixed", synth.alarm37(), dict(exact_k=4), 0.0, 20
    yield "alarm37_soft", synth.alarm37(), dict(exact_k=4, soft=True), 0.0, 15
    yield "polytree40", synth.random_polytree(40, card_hi=4, seed=21), dict(p=0.15), 1e-8, 300
    yield "grid6_loopy", synth.grid(6, seed=4), dict(p=0.1), 1e-7, 400
    yield "dag45_k4", synth.random_dag(45, 4, 2, 3, seed=7), dict(p=0.1), 0.0, 12
    yield "hub12_isolated", _hub(), dict(p=0.3), 1e-9, 200


@pytest.mark.parametrize("name,net,evkw,eps,cap", list(_cases()), ids=[c[0] for c in _cases()])
@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_specialised_vs_oracle(BP, oracle_mod, name, net, evkw, eps, cap, precision):
    if precision == "fp32" and eps > 0:
        pytest.skip("fp32 parity is asserted at fixed sweep counts")
    ev = synth.make_evidence(net, 777, seed=17, **evkw)          # ragged: not a multiple of any tile
    om, osw, ocv = oracle_mod.run_port(net, ev, eps=eps, max_sweeps=cap, threads=0)
    bp = BP(net, precision, specialize="always")
    res = bp(ev, eps, max_sweeps=cap)
    assert bp.stats()["last_specialised"] == 1
    assert np.array_equal(res.sweeps, osw), (name, np.nonzero(res.sweeps != osw)[0][:5])
    assert np.array_equal(res.converged, ocv), name
    assert_close(res.marginals, om, what=name, **TOL[precision])


@pytest.mark.parametrize("name", ["pearl_tests", "pearl_nan", "pearl_nan_fixed6", "resume_tests", "resume_soft"])
def test_specialised_reference_fixtures(BP, ref_fixtures, name):
    """The reference's own graphs incl. the impossible-evidence (NaN) case, from reference output."""
    f = load_fixture(ref_fixtures, name)
    bp = BP(f["net"], "fp64", specialize="always")
    res = bp(f["ev"], f["eps"], max_sweeps=f["max_sweeps"])
    assert bp.stats()["last_specialised"] == 1
    assert np.array_equal(res.sweeps, f["sweeps"]) and np.array_equal(res.converged, f["converged"])
    assert_close(res.marginals, f["marginals"], what=name, **TOL["fp64"])


def test_specialised_extensions_and_refresh(BP, oracle_mod):
    """damping / check_interval (variants 1 and 2) and bnbp_refresh_cpt into the constant bank."""
    net = synth.grid(5, seed=8)
    ev = synth.make_evidence(net, 300, p=0.15, seed=5)
    bp = BP(net, specialize="always")
    for kw in (dict(damping=0.25), dict(check_interval=4), dict(damping=0.1, check_interval=3)):
        om, osw, ocv = oracle_mod.run_port(net, ev, eps=1e-8, max_sweeps=500, threads=0, **kw)
        res = bp(ev, 1e-8, max_sweeps=500, **kw)
        assert bp.stats()["last_specialised"] == 1
        assert np.array_equal(res.sweeps, osw) and np.array_equal(res.converged, ocv), kw
        assert_close(res.marginals, om, what=str(kw), **TOL["fp64"])
    net2 = synth.grid(5, seed=9)                                  # same topology, other CPTs
    bp.refresh_cpt(net2.cpt)
    om, _, _ = oracle_mod.run_port(net2, ev, eps=0.0, max_sweeps=9)
    assert_close(bp(ev, 0.0, max_sweeps=9).marginals, om, what="refresh", **TOL["fp64"])


@pytest.mark.parametrize("cap", [1, 2, 3])
def test_specialised_short_fixed_runs(BP, oracle_mod, cap):
    """1, 2 and 3 fixed sweeps: the first sweep (no message loads) and the last sweep (no message
    stores) are separate kernel variants; 1 sweep uses the plain variant alone."""
    net = synth.grid(5, seed=8)
    ev = synth.make_evidence(net, 300, p=0.15, seed=5)
    om, osw, _ = oracle_mod.run_port(net, ev, eps=0.0, max_sweeps=cap)
    res = BP(net, specialize="always")(ev, 0.0, max_sweeps=cap)
    assert np.array_equal(res.sweeps, osw)
    assert_close(res.marginals, om, what=f"{cap} sweeps", **TOL["fp64"])


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("n_cases", [1, 777, 4096 + 33])
def test_fused_first_and_last_sweep_bitwise(BP, monkeypatch, precision, n_cases):
    """Fixed sweeps + hard evidence: K0 runs inside the first sweep and K4 inside the last (variants 5
    and 6/7 of bnbp_spec.cuh).  Same arithmetic in the same order as the unfused launch sequence, so
    the marginals are bit-identical -- through the host-buffer call (doubles out of a float kernel:
    variant 7) and through the device call (marginals in the kernel's own type: variant 6)."""
    import torch
    net = synth.alarm37()
    ev = synth.make_evidence(net, n_cases, exact_k=4, seed=3)
    bp = BP(net, precision, specialize="always")
    dev = torch.device("cuda", 0)
    d_off, d_node, d_state = (torch.from_numpy(a).to(dev) for a in (ev.ev_off, ev.ev_node, ev.ev_state))
    tdt = torch.float64 if precision == "fp64" else torch.float32

    def both():
        host = bp(ev, 0.0, max_sweeps=6)
        fused = bp.stats()["last_fused"]
        d_out = torch.full((n_cases, net.belief_values), -7.0, dtype=tdt, device=dev)
        d_sw = torch.zeros(n_cases, dtype=torch.int32, device=dev)
        bp.run_device(n_cases, d_off, d_node, d_state, d_out, epsilon=0.0, max_sweeps=6, out_sweeps=d_sw)
        torch.cuda.synchronize()
        return host, d_out.cpu().numpy(), d_sw.cpu().numpy(), fused, bp.stats()["last_fused"]

    h1, dv1, sw1, f1a, f1b = both()
    assert f1a == 1 and f1b == 1
    monkeypatch.setenv("BNBP_NO_FUSE", "1")
    h0, dv0, sw0, f0a, f0b = both()
    assert f0a == 0 and f0b == 0
    assert np.array_equal(h1.marginals, h0.marginals, equal_nan=True)
    assert np.array_equal(dv1, dv0, equal_nan=True)
    assert np.array_equal(h1.sweeps, h0.sweeps) and np.array_equal(sw1, sw0) and np.all(sw1 == 6)
    assert np.array_equal(h1.converged, h0.converged)


def test_fused_sequence_soft_evidence_falls_back_to_k0_k4(BP, oracle_mod):
    """Soft evidence rows do not fit the one-byte state table of the fused first sweep: the unfused
    sequence (still the specialised GPU kernels) runs, and matches the oracle."""
    net = synth.alarm37()
    ev = synth.make_evidence(net, 500, exact_k=4, soft=True, seed=5)
    bp = BP(net, "fp64", specialize="always")
    res = bp(ev, 0.0, max_sweeps=8)
    st = bp.stats()
    assert st["last_specialised"] == 1 and st["last_fused"] == 0
    om, _, _ = oracle_mod.run_port(net, ev, eps=0.0, max_sweeps=8, threads=0)
    assert_close(res.marginals, om, what="soft", **TOL["fp64"])


@pytest.mark.parametrize("family", ["always", "never"])
@pytest.mark.parametrize("kw", [dict(), dict(damping=0.2), dict(check_interval=3)], ids=["plain", "damped", "interval3"])
def test_eps_mode_compaction_of_active_cases(BP, oracle_mod, monkeypatch, family, kw):
    """eps mode on a batch whose sweep counts spread: the converged cases are retired and the active ones
    gathered into dense tiles at checkpoints (compact_* kernels).  A case's arithmetic does not depend on
    its position, so marginals, sweep counts and flags are bit-identical to the run without compaction,
    and match the oracle."""
    net = synth.alarm37()
    n = 20000 + 77
    ev = synth.make_evidence(net, n, exact_k=4, seed=13)
    bp = BP(net, "fp64", specialize=family)
    a = bp(ev, 1e-6, max_sweeps=200, **kw)
    st = bp.stats()
    assert st["last_compactions"] >= 1, st
    monkeypatch.setenv("BNBP_NO_COMPACT", "1")
    b = bp(ev, 1e-6, max_sweeps=200, **kw)
    assert bp.stats()["last_compactions"] == 0
    assert np.array_equal(a.sweeps, b.sweeps) and np.array_equal(a.converged, b.converged)
    assert np.array_equal(a.marginals, b.marginals, equal_nan=True)
    # the specialised family without damping runs plain sweeps + delta_retire_kernel ("split"); the
    # freeze/check variants of the sweep kernel (BNBP_NO_SPLIT) must give the same bits
    monkeypatch.setenv("BNBP_NO_SPLIT", "1")
    c = bp(ev, 1e-6, max_sweeps=200, **kw)
    assert np.array_equal(a.sweeps, c.sweeps) and np.array_equal(a.converged, c.converged)
    if family == "always":
        assert np.array_equal(a.marginals, c.marginals, equal_nan=True)
    else:
        # the generic kernel's plain and freeze/check forms are separate template instantiations: nvcc
        # contracts a few multiply-adds differently, so they agree to rounding, not to the bit
        assert_close(a.marginals, c.marginals, rtol=1e-12, atol=1e-15, what="generic split vs in-kernel test")
    om, osw, ocv = oracle_mod.run_port(net, ev, eps=1e-6, max_sweeps=200, threads=0, **kw)
    assert np.array_equal(a.sweeps, osw) and np.array_equal(a.converged, ocv)
    assert_close(a.marginals, om, what="compaction " + family, **TOL["fp64"])
    assert a.sweeps.max() > 2 * a.sweeps.min()                 # the spread that makes compaction pay


def test_split_eps_path_on_impossible_evidence(BP, oracle_mod, ref_fixtures):
    """The reference's Pearl graph with impossible evidence (fixture pearl_nan_fixed6: NaN beliefs after 6
    fixed sweeps), in a batch large enough for the split eps path.  Under the reference's stopping rule these
    cases stop after 2-3 sweeps, before the 0/0 appears; the split path must stop them at the same sweep."""
    f = load_fixture(ref_fixtures, "pearl_nan_fixed6")
    net, ev = f["net"], f["ev"]
    reps = 20000 // ev.n_cases + 1
    off = np.concatenate([[0], np.cumsum(np.tile(np.diff(ev.ev_off), reps))]).astype(np.int64)
    big = EvidenceBatch(ev.n_cases * reps, off, np.tile(ev.ev_node, reps), np.tile(ev.ev_state, reps))
    bp = BP(net, "fp64", specialize="always")
    for eps in (1e-3, 1e-9):
        res = bp(big, eps, max_sweeps=50)
        om, osw, ocv = oracle_mod.run_port(net, big, eps=eps, max_sweeps=50, threads=0)
        assert np.array_equal(res.sweeps, osw) and np.array_equal(res.converged, ocv)
        assert_close(res.marginals, om, what="impossible evidence, eps mode", **TOL["fp64"])


def test_eps_mode_compaction_with_sweep_cap_and_fp32(BP, oracle_mod):
    """Cases that hit max_sweeps without converging stay in the arena to the end; fp32 handle, device call."""
    import torch
    net = synth.alarm37()
    n = 40000
    ev = synth.make_evidence(net, n, exact_k=4, seed=14)
    bp = BP(net, "fp64", specialize="always")
    res = bp(ev, 1e-9, max_sweeps=14)                          # most cases do not make 1e-9 in 14 sweeps
    om, osw, ocv = oracle_mod.run_port(net, ev, eps=1e-9, max_sweeps=14, threads=0)
    assert np.array_equal(res.sweeps, osw) and np.array_equal(res.converged, ocv)
    assert 0 < int(ocv.sum()) < n
    assert_close(res.marginals, om, what="capped", **TOL["fp64"])
    bp32 = BP(net, "fp32", specialize="always")
    dev = torch.device("cuda", 0)
    d_off, d_node, d_state = (torch.from_numpy(a).to(dev) for a in (ev.ev_off, ev.ev_node, ev.ev_state))
    d_out = torch.zeros((n, net.belief_values), dtype=torch.float32, device=dev)
    d_sw = torch.zeros(n, dtype=torch.int32, device=dev)
    bp32.run_device(n, d_off, d_node, d_state, d_out, epsilon=1e-3, max_sweeps=100, out_sweeps=d_sw)
    torch.cuda.synchronize()
    assert bp32.stats()["last_compactions"] >= 1
    ref, rsw, _ = oracle_mod.run_port(net, ev, eps=1e-3, max_sweeps=100, threads=0)
    # fp32 sweep counts may differ by one where delta sits at the threshold; the marginals agree to fp32 accuracy
    assert np.mean(d_sw.cpu().numpy() == rsw) > 0.99
    assert np.abs(d_out.cpu().numpy() - ref).max() < 5e-3


def test_specialised_equals_generic_bitwise_shape(BP):
    """Both kernel families implement the same schedule: fixed sweeps, same sweep counts, results
    within a few ulp of each other (they differ only in the order of some products)."""
    net = synth.alarm37()
    ev = synth.make_evidence(net, 5000, exact_k=4)
    a = BP(net, specialize="always")(ev, 0.0, max_sweeps=20)
    b = BP(net, specialize="never")(ev, 0.0, max_sweeps=20)
    assert_close(a.marginals, b.marginals, rtol=1e-12, atol=1e-15, what="spec vs generic")


def test_ineligible_network_is_refused_when_forced(BP):
    from bayesiannetwork_b200.engine import BnbpError
    net = synth.high_card(6, card=32, n_parents=2, seed=10)       # cardinality 32: not specialisable
    bp = BP(net, specialize="always")
    with pytest.raises(BnbpError):
        bp(EvidenceBatch.empty(4), 0.0, max_sweeps=2)
    assert BP(net, specialize="auto")(EvidenceBatch.empty(4), 0.0, max_sweeps=2).marginals.shape[0] == 4
