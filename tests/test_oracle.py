"""Pins the oracle (CPU restatement, oracle/bp_oracle.c) -- runs without a GPU.

1. against the reference's own test vectors (tests/golden/reference_tests.json),
2. against fixtures produced by the reference itself (tests/golden/ref_fixtures.npz),
3. live against the compiled reference (oracle/_ref) when it is present.
"""
import json
import os

import numpy as np
import pytest

from bayesiannetwork_b200 import synth
from bayesiannetwork_b200.flat import EvidenceBatch
from helpers import assert_close, load_fixture

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SPEC = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_tests.json")))
NETS = {"pearl": synth.pearl_network, "resume": synth.resume_network}


def boost_check_close(a, b, tol_percent):
    """BOOST_CHECK_CLOSE: strong check, both relative differences within tol (percent)."""
    if a == b:
        return True
    d = abs(a - b)
    return d <= tol_percent / 100 * abs(a) and d <= tol_percent / 100 * abs(b)


@pytest.mark.parametrize("case", SPEC["cases"], ids=[c["name"] for c in SPEC["cases"]])
def test_reference_test_vectors(oracle_mod, case):
    net = NETS[case["network"]]()
    ev = EvidenceBatch.from_cases(net, [{int(k): v for k, v in case["evidence"].items()}])
    m, sweeps, conv = oracle_mod.run_port(net, ev, eps=case["eps"])
    off = net.belief_off
    assert sweeps[0] == case["sweeps"] and conv[0] == 1
    for node, teacher in case["teacher"].items():
        got = m[0, off[int(node)]:off[int(node) + 1]]
        for g, t in zip(got, teacher):
            assert boost_check_close(float(g), t, case["tol_percent"]), (case["name"], node, got, teacher)
    if case["beliefs"] is not None:
        flat = np.concatenate([np.asarray(b, dtype=np.float64) for b in case["beliefs"]])
        assert_close(m[0], flat, rtol=1e-12, atol=1e-15, what=case["name"])


def test_fixture_names(ref_fixtures):
    assert len(ref_fixtures["names"]) >= 15


@pytest.mark.parametrize("name", [
    "pearl_tests", "resume_tests", "resume_soft", "pearl_nan", "pearl_nan_fixed6", "pearl_one_sweep",
    "polytree24_eps", "polytree24_soft", "grid4_eps", "grid4_fixed7", "grid6_fixed12", "dag40_eps",
    "dag40_fixed5", "alarm37_eps", "alarm37_fixed20", "card6_fixed6"])
def test_port_matches_reference_fixtures(oracle_mod, ref_fixtures, name):
    f = load_fixture(ref_fixtures, name)
    m, sweeps, conv = oracle_mod.run_port(f["net"], f["ev"], eps=f["eps"], max_sweeps=f["max_sweeps"])
    assert np.array_equal(sweeps, f["sweeps"]), name
    assert np.array_equal(conv, f["converged"]), name
    assert_close(m, f["marginals"], rtol=1e-12, atol=1e-15, what=name)


def test_nan_fixture_really_has_nan(ref_fixtures):
    f = load_fixture(ref_fixtures, "pearl_nan_fixed6")
    assert np.isnan(f["marginals"]).any()


def test_port_threads_agree(oracle_mod):
    net = synth.alarm37()
    ev = synth.make_evidence(net, 64, exact_k=4)
    a = oracle_mod.run_port(net, ev, eps=1e-6, max_sweeps=100, threads=1)
    b = oracle_mod.run_port(net, ev, eps=1e-6, max_sweeps=100, threads=4)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_port_extensions_default_off(oracle_mod):
    """damping=0 / check_interval=1 are the reference behaviour; check_interval only rounds the
    stopping sweep up to a multiple."""
    net = synth.grid(4, seed=5)
    ev = synth.make_evidence(net, 8, p=0.2, seed=3)
    m1, s1, _ = oracle_mod.run_port(net, ev, eps=1e-8, max_sweeps=400)
    m4, s4, c4 = oracle_mod.run_port(net, ev, eps=1e-8, max_sweeps=400, check_interval=4)
    assert np.all(s4 % 4 == 0) and np.all(s4 >= s1) and np.all(s4 < s1 + 4) and c4.all()
    md, sd, cd = oracle_mod.run_port(net, ev, eps=1e-8, max_sweeps=400, damping=0.3)
    assert cd.all()
    assert_close(md, m1, rtol=1e-5, atol=1e-7, what="damped fixed point")


def _have_ref(oracle_mod):
    return oracle_mod.have_reference() and os.path.exists(oracle_mod.REF_PURE_SO)


def test_live_reference_capped_equals_pure(oracle_mod):
    """The sweep-cap shim does not change the reference's results (eps mode), and eps=1e300
    means exactly one sweep."""
    if not _have_ref(oracle_mod):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    net = synth.random_dag(30, max_parents=3, card_lo=2, card_hi=4, seed=5)
    ev = synth.make_evidence(net, 6, p=0.15, seed=2)
    mc, sc, cc, _ = oracle_mod.run_reference(net, ev, eps=1e-6, max_sweeps=1000)
    mp, _, _, _ = oracle_mod.run_reference(net, ev, eps=1e-6, pure=True)
    # not bit-equal: the reference multiplies in pointer-keyed unordered_map order
    assert_close(mc, mp, rtol=1e-13, atol=1e-16, what="capped vs pure")
    assert cc.all()
    m1, s1, _, _ = oracle_mod.run_reference(net, ev, eps=1e300, max_sweeps=1000)
    assert np.all(s1 == 1)
    mport, sport, _ = oracle_mod.run_port(net, ev, eps=1e-6, max_sweeps=1000)
    assert np.array_equal(sport, sc)
    assert_close(mport, mc, rtol=1e-12, atol=1e-15, what="live")


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_live_reference_random_nets(oracle_mod, seed):
    if not _have_ref(oracle_mod):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    net = synth.random_dag(25 + 5 * seed, max_parents=4, card_lo=2, card_hi=5, seed=seed)
    for soft in (False, True):
        ev = synth.make_evidence(net, 5, p=0.2, seed=seed, soft=soft)
        for eps, cap in ((1e-7, 300), (0.0, 9)):
            mr, sr, cr, _ = oracle_mod.run_reference(net, ev, eps=eps, max_sweeps=cap)
            mp, sp, cp = oracle_mod.run_port(net, ev, eps=eps, max_sweeps=cap)
            assert np.array_equal(sr, sp) and np.array_equal(cr, cp)
            assert_close(mp, mr, rtol=1e-12, atol=1e-15, what=f"seed{seed} soft{soft} eps{eps}")


def test_torch_evidence_generator_is_bit_identical_to_numpy():
    """synth.make_evidence_torch (what bench.py uses for the big batches of configs 3-5) = synth.make_evidence."""
    from bayesiannetwork_b200 import synth
    for net in (synth.grid(9), synth.random_dag(120), synth.high_card(6, 32, 3)):
        for off in (0, 98765):
            a = synth.make_evidence(net, 300, p=0.1, case_offset=off)
            o, nd, st = synth.make_evidence_torch(net, 300, p=0.1, case_offset=off, device="cpu", chunk_elems=20000)
            assert np.array_equal(a.ev_off, o.numpy()) and np.array_equal(a.ev_node, nd.numpy())
            assert np.array_equal(a.ev_state, st.numpy())


def _max_marginals_brute_force(net, evidence):
    """max over all completions of P(x, rest) for every (node, state): what max-product BP computes on a polytree."""
    import itertools
    best = [np.zeros(int(r)) for r in net.card]
    for st in itertools.product(*[range(int(r)) for r in net.card]):
        if any(st[n] != s for n, s in evidence.items()):
            continue
        p = 1.0
        for x in range(net.n_nodes):
            q = 0
            for u in net.parents[net.parent_off[x]:net.parent_off[x + 1]]:
                q = q * int(net.card[u]) + st[u]
            p *= net.cpt[net.cpt_off[x] + q * int(net.card[x]) + st[x]]
        for x in range(net.n_nodes):
            best[x][st[x]] = max(best[x][st[x]], p)
    return np.concatenate([b / b.sum() for b in best])


def test_max_product_port_gives_max_marginals_on_polytrees(oracle_mod):
    """The oracle's max-product switch (SURVEY 8 f4 extension; the reference is sum-product only): on a polytree the
    normalised beliefs are the normalised max-marginals -- pinned by brute-force enumeration."""
    net = synth.random_polytree(9, card_hi=3, max_parents=2, seed=5)
    cases = [{}, {2: 1}, {0: 0, 7: 1}, {4: 0, 5: 1, 8: 0}]
    cases = [{n: s % int(net.card[n]) for n, s in c.items()} for c in cases]
    ev = EvidenceBatch.from_cases(net, cases)
    got, sw, cv = oracle_mod.run_port(net, ev, eps=1e-13, max_sweeps=100, semiring=1)
    assert cv.all()
    for c, case in enumerate(cases):
        assert_close(got[c], _max_marginals_brute_force(net, case), rtol=1e-10, atol=1e-13, what=f"max-marginals case {c}")
    # and it is a different thing from the sum-product marginals
    sump, _, _ = oracle_mod.run_port(net, ev, eps=1e-13, max_sweeps=100)
    assert np.abs(sump - got).max() > 1e-3


def test_textbook_burglary_network_known_answer(oracle_mod):
    """A number from the literature, independent of this repo and of the reference: Pearl's burglary / earthquake alarm
    network with the probabilities of Russell & Norvig (AIMA, fig. 14.2) gives P(Burglary | JohnCalls, MaryCalls) = 0.284
    (AIMA section 14.4).  The network is a polytree, so belief propagation is exact on it."""
    # nodes: 0 Burglary, 1 Earthquake, 2 Alarm (parents 0, 1), 3 JohnCalls (2), 4 MaryCalls (2); state 0 = true, 1 = false
    card = [2, 2, 2, 2, 2]
    parents = [[], [], [0, 1], [2], [2]]
    net = synth._assemble(card, parents, 1, "burglary")
    cpt = [0.001, 0.999,                      # P(B)
           0.002, 0.998,                      # P(E)
           0.95, 0.05, 0.94, 0.06, 0.29, 0.71, 0.001, 0.999,    # P(A | B, E): rows (t,t) (t,f) (f,t) (f,f), first parent slowest
           0.90, 0.10, 0.05, 0.95,            # P(J | A)
           0.70, 0.30, 0.01, 0.99]            # P(M | A)
    net.cpt[:] = cpt
    ev = EvidenceBatch.from_cases(net, [{3: 0, 4: 0}, {}])
    m, sw, cv = oracle_mod.run_port(net, ev, eps=1e-12, max_sweeps=100)
    assert cv.all()
    assert abs(m[0, 0] - 0.284) < 5e-4                        # P(b | j, m) = 0.284
    exact = 0.001 * (0.002 * 0.95 * 0.9 * 0.7 + 0.998 * 0.94 * 0.9 * 0.7 + 0.002 * 0.05 * 0.05 * 0.01 + 0.998 * 0.06 * 0.05 * 0.01)
    other = 0.999 * (0.002 * 0.29 * 0.9 * 0.7 + 0.998 * 0.001 * 0.9 * 0.7 + 0.002 * 0.71 * 0.05 * 0.01 + 0.998 * 0.999 * 0.05 * 0.01)
    assert abs(m[0, 0] - exact / (exact + other)) < 1e-12
    assert abs(m[1, 0] - 0.001) < 1e-15 and abs(m[1, 4 * 2 - 2] - 0.05210) < 5e-5          # priors: P(b), P(j) = 0.0521
