"""Likelihood weighting and CPT estimation from samples.

CPU part (``-m "not gpu"``): the C restatements in oracle/bp_oracle.c against exact enumeration,
against the reference's own likelihood_weighting.hpp compiled in place (statistically: the reference
seeds from std::random_device) and against hand-computed CPT tables.
GPU part (``-m gpu``): the CUDA kernels behind bnbp_lw_run_batch / bnbp_estimate_cpt against those
restatements -- draw for draw (same counter-based variates) and count for count."""
import itertools
import os

import numpy as np
import pytest

from bayesiannetwork_b200 import synth
from bayesiannetwork_b200.flat import EvidenceBatch, FlatNetwork
from helpers import assert_close

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def exact_posteriors(net: FlatNetwork, ev: EvidenceBatch) -> np.ndarray:
    """Brute-force P(X | evidence) for every node by enumerating the joint (small networks only)."""
    n, card, boff = net.n_nodes, net.card, net.belief_off
    states = np.array(list(itertools.product(*[range(int(c)) for c in card])), dtype=np.int64)
    joint = np.ones(states.shape[0])
    for x in range(n):
        q = np.zeros(states.shape[0], dtype=np.int64)
        for e in range(net.parent_off[x], net.parent_off[x + 1]):
            p = net.parents[e]
            q = q * card[p] + states[:, p]
        joint *= net.cpt[net.cpt_off[x] + q * card[x] + states[:, x]]
    out = np.zeros((ev.n_cases, net.belief_values))
    for c in range(ev.n_cases):
        w = joint.copy()
        for e in range(ev.ev_off[c], ev.ev_off[c + 1]):
            w *= states[:, ev.ev_node[e]] == ev.ev_state[e]
        for x in range(n):
            for s in range(card[x]):
                out[c, boff[x] + s] = w[states[:, x] == s].sum()
            out[c, boff[x]:boff[x + 1]] /= w.sum()
    return out


def ancestral_samples(net: FlatNetwork, n_rows: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    s = np.zeros((n_rows, net.n_nodes), dtype=np.int32)
    for x in range(net.n_nodes):                     # synth networks: parents have smaller ids
        q = np.zeros(n_rows, dtype=np.int64)
        for e in range(net.parent_off[x], net.parent_off[x + 1]):
            p = net.parents[e]
            q = q * net.card[p] + s[:, p]
        rows = net.cpt[net.cpt_off[x] + q[:, None] * net.card[x] + np.arange(net.card[x])[None, :]]
        u = rng.random(n_rows)[:, None]
        s[:, x] = np.minimum((np.cumsum(rows, axis=1) <= u).sum(axis=1), net.card[x] - 1)
    return s


# ---- CPU: the restatements ---------------------------------------------------------------------------
def test_lw_port_matches_exact_enumeration(oracle_mod):
    net = synth.random_dag(9, 3, 2, 3, seed=5)                  # loopy: BP is approximate here, LW is not
    ev = synth.make_evidence(net, 6, exact_k=2, seed=3)
    want = exact_posteriors(net, ev)
    got, wsum = oracle_mod.run_port_lw(net, ev, 200000, seed=11)
    assert np.all(wsum > 0)
    assert np.abs(got - want).max() < 0.01                      # ~5 sigma of a 0.5 proportion at 2e5 samples / weight spread
    boff = net.belief_off
    for x in range(net.n_nodes):
        assert np.allclose(got[:, boff[x]:boff[x + 1]].sum(axis=1), 1.0, atol=1e-12)


def test_lw_port_reproducible_and_case_keyed(oracle_mod):
    net = synth.alarm37()
    ev = synth.make_evidence(net, 5, exact_k=4, seed=2)
    a, _ = oracle_mod.run_port_lw(net, ev, 500, seed=7)
    b, _ = oracle_mod.run_port_lw(net, ev, 500, seed=7)
    assert np.array_equal(a, b)
    c, _ = oracle_mod.run_port_lw(net, ev.slice(2, 5), 500, seed=7, case_base=2)   # a shard sees the same variates
    assert np.array_equal(a[2:5], c)
    d, _ = oracle_mod.run_port_lw(net, ev, 500, seed=8)
    assert not np.array_equal(a, d)


def test_lw_port_vs_reference_statistically(oracle_mod):
    if not oracle_mod.have_reference_lw():
        pytest.skip("oracle/_ref/libbnref_lw.so not built (needs /root/reference)")
    net = synth.random_dag(10, 3, 2, 4, seed=9)
    ev = synth.make_evidence(net, 3, exact_k=2, seed=4)
    ref = oracle_mod.run_reference_lw(net, ev, 60000)
    got, _ = oracle_mod.run_port_lw(net, ev, 60000, seed=5)
    assert np.abs(ref - got).max() < 0.03
    assert np.abs(ref - exact_posteriors(net, ev)).max() < 0.03


def test_lw_impossible_evidence_gives_uniform_rows(oracle_mod):
    """All samples weigh 0 -> normalize() returns uniform rows (likelihood_weighting.hpp:206-213)."""
    card = [2, 2]
    net = synth._assemble(card, [[], [0]], 1, "det")
    cpt = net.cpt.copy()
    cpt[net.cpt_off[1]:net.cpt_off[2]] = [1.0, 0.0, 1.0, 0.0]     # node 1 is always 0
    net = FlatNetwork(net.card, net.parent_off, net.parents, net.cpt_off, cpt)
    ev = EvidenceBatch.from_cases(net, [{1: 1}])
    got, wsum = oracle_mod.run_port_lw(net, ev, 100, seed=1)
    assert wsum[0] == 0.0 and np.array_equal(got[0], [0.5, 0.5, 0.5, 0.5])


def test_make_cpt_port_known_answer(oracle_mod):
    """A -> B, both binary.  4 distinct samples with multiplicities (the reference's file format:
    count, then one state per node; sampler.hpp:53-76)."""
    net = synth._assemble([2, 2], [[], [0]], 1, "ab")
    samples = np.array([[0, 0], [0, 1], [1, 1], [0, 0]], dtype=np.int32)
    mult = np.array([3, 1, 4, 2], dtype=np.int64)
    cpt = oracle_mod.port_make_cpt(net, samples, mult)
    # A: (3+1+2, 4)/10 ; B|A=0: (5, 1)/6 ; B|A=1: (0, 4)/4
    assert np.array_equal(cpt, np.array([0.6, 0.4, 5 / 6, 1 / 6, 0.0, 1.0]))
    # a parent configuration that never occurs -> uniform row (sampler.hpp:147-151)
    cpt2 = oracle_mod.port_make_cpt(net, samples[:2], mult[:2])
    assert np.array_equal(cpt2, np.array([1.0, 0.0, 0.75, 0.25, 0.5, 0.5]))
    # multiplicity m == the row repeated m times
    rep = np.repeat(samples, mult, axis=0)
    assert np.array_equal(oracle_mod.port_make_cpt(net, rep), cpt)


def test_make_cpt_port_recovers_the_generating_network(oracle_mod):
    net = synth.random_dag(12, 3, 2, 4, seed=3)
    s = ancestral_samples(net, 400000, seed=1)
    est = oracle_mod.port_make_cpt(net, s)
    # rows of rarely visited parent configurations are noisy: weigh the error by sqrt(visits)
    assert np.abs(est - net.cpt).max() < 0.08
    assert np.median(np.abs(est - net.cpt)) < 0.004


# ---- GPU: the kernels ---------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def BP():
    from bayesiannetwork_b200.engine import BeliefPropagation
    return BeliefPropagation


@pytest.mark.gpu
@pytest.mark.parametrize("name,net,evkw,n_samples", [
    ("alarm37", synth.alarm37(), dict(exact_k=4), 3000),
    ("dag30_loopy", synth.random_dag(30, 4, 2, 5, seed=12), dict(p=0.15), 1000),
    ("grid6", synth.grid(6, seed=4), dict(p=0.1), 777),
], ids=["alarm37", "dag30_loopy", "grid6"])
def test_gpu_lw_matches_port_draw_for_draw(BP, oracle_mod, name, net, evkw, n_samples):
    ev = synth.make_evidence(net, 70, seed=23, **evkw)
    want, wwant = oracle_mod.run_port_lw(net, ev, n_samples, seed=99)
    got, wgot = BP(net).likelihood_weighting(ev, n_samples, seed=99, return_weight=True)
    # same variates, same selections; only the order of the fp64 weight sums differs (shared-memory atomics)
    assert_close(got, want, rtol=1e-11, atol=1e-13, what=name)
    assert_close(wgot, wwant, rtol=1e-11, atol=0.0, what=name + " weight")


@pytest.mark.gpu
def test_gpu_lw_agrees_with_bp_on_a_polytree(BP):
    """BP is exact on polytrees, so LW must land on the BP marginals within sampling error."""
    net = synth.random_polytree(25, card_hi=3, seed=6)
    ev = synth.make_evidence(net, 16, exact_k=2, seed=8)
    bp = BP(net)
    exact = bp(ev, 1e-10, max_sweeps=200).marginals
    lw = bp.likelihood_weighting(ev, 200000, seed=3)
    assert np.abs(lw - exact).max() < 0.02


@pytest.mark.gpu
def test_gpu_lw_rejects_soft_and_bad_evidence(BP):
    from bayesiannetwork_b200.engine import BnbpError
    net = synth.alarm37()
    bp = BP(net)
    with pytest.raises(ValueError):
        bp.likelihood_weighting(synth.make_evidence(net, 4, exact_k=2, soft=True), 10)
    bad = EvidenceBatch(1, np.array([0, 1], dtype=np.int64), np.array([3], dtype=np.int32), np.array([99], dtype=np.int32))
    with pytest.raises(BnbpError):
        bp.likelihood_weighting(bad, 10)


@pytest.mark.gpu
@pytest.mark.parametrize("with_mult", [False, True])
def test_gpu_estimate_cpt_equals_port_exactly(oracle_mod, with_mult):
    from bayesiannetwork_b200.engine import estimate_cpt
    net = synth.random_dag(40, 4, 2, 6, seed=31)
    s = ancestral_samples(net, 50000, seed=2)
    mult = (np.arange(s.shape[0]) % 7 + 1).astype(np.int64) if with_mult else None
    want = oracle_mod.port_make_cpt(net, s, mult)
    got = estimate_cpt(net, s, mult)
    assert np.array_equal(got, want)                            # integer counts: no rounding freedom


@pytest.mark.gpu
def test_gpu_estimate_cpt_then_infer(BP, oracle_mod):
    """The reference workflow sampler::make_cpt -> belief_propagation, without leaving the library."""
    from bayesiannetwork_b200.engine import BnbpError, estimate_cpt
    net = synth.alarm37()
    s = ancestral_samples(net, 200000, seed=4)
    est = FlatNetwork(net.card, net.parent_off, net.parents, net.cpt_off, estimate_cpt(net, s))
    ev = synth.make_evidence(net, 64, exact_k=4, seed=6)
    a = BP(net)(ev, 0.0, max_sweeps=20).marginals
    b = BP(est)(ev, 0.0, max_sweeps=20).marginals
    assert np.abs(a - b).max() < 0.08
    s[5, 3] = 77
    with pytest.raises(BnbpError):
        estimate_cpt(net, s)


# ---- f3 pinned against the reference's own sampler.hpp (compiled in place over Boost stand-ins) -----------------
SAMPLER_FIXTURES = ["pearl", "alarm37", "dag25_card5"]


def _sampler_fixture(name):
    fx = np.load(os.path.join(ROOT, "tests", "golden", "sampler_fixture.npz"), allow_pickle=False)
    p = name + "/"
    net = FlatNetwork(fx[p + "card"], fx[p + "parent_off"], fx[p + "parents"], fx[p + "cpt_off"],
                      np.zeros(int(fx[p + "cpt_off"][-1])), name=name)
    return net, fx[p + "samples"], fx[p + "mult"], fx[p + "cpt"]


@pytest.mark.parametrize("name", SAMPLER_FIXTURES)
def test_make_cpt_port_equals_reference_golden(oracle_mod, name):
    """oracle/bp_oracle.c::bp_oracle_make_cpt == the reference's sampler::make_cpt (sampler.hpp:81-163), from the
    committed output of the reference binary: counts are integers and count / total is one correctly rounded
    division in both, so the rows agree bit for bit (bar 1e-15)."""
    net, s, mult, want = _sampler_fixture(name)
    got = oracle_mod.port_make_cpt(net, s, mult)
    assert np.abs(got - want).max() <= 1e-15 and np.array_equal(got, want)
    assert (want.reshape(-1) == 0).sum() >= 0 and np.isfinite(want).all()


def test_make_cpt_port_equals_live_reference_and_its_file_reader(oracle_mod, tmp_path):
    """Where the reference could be compiled (this container): live differential run incl. the reference's own
    sample-file reader (sampler.hpp:42-76) on a file with runs of blanks (token_compress_on)."""
    if not oracle_mod.have_reference_sampler():
        pytest.skip("oracle/_ref/libbnref_sampler.so not built (needs /root/reference)")
    net = synth.random_dag(18, 3, 2, 4, seed=9)
    s = np.unique(ancestral_samples(net, 5000, seed=3), axis=0)
    mult = (np.arange(s.shape[0]) % 5 + 1).astype(np.int64)
    want = oracle_mod.reference_make_cpt(net, s, mult)
    assert np.array_equal(oracle_mod.port_make_cpt(net, s, mult), want)
    path = tmp_path / "samples.txt"
    path.write_text("".join(f"{int(m)}  " + " ".join(str(int(v)) for v in row) + "\n" for row, m in zip(s, mult)))
    from_file, size = oracle_mod.reference_make_cpt_from_file(net, str(path))
    assert size == int(mult.sum()) and np.array_equal(from_file, want)


@pytest.mark.gpu
@pytest.mark.parametrize("name", SAMPLER_FIXTURES)
def test_gpu_estimate_cpt_equals_reference_golden(name):
    """bnbp_estimate_cpt (cpt_count_kernel + cpt_normalize_kernel) against the reference binary's output."""
    from bayesiannetwork_b200.engine import estimate_cpt
    net, s, mult, want = _sampler_fixture(name)
    got = estimate_cpt(net, s, mult)
    assert np.abs(got - want).max() <= 1e-15 and np.array_equal(got, want)
