"""Network files: BIF / DSC text -> flat network through the C ABI
(bnbp_netfile_*, i.e. include/bayesian/serializer/{bif,dsc}.hpp).  CPU only.

Pinning: the DSC loader is compared with the reference's OWN DSC loader (serializer/dsc.hpp needs no
Boost: oracle/_ref/libbnref_dsc.so, live where /root/reference exists, and the committed fixture
tests/golden/dsc_fixture.npz it produced).  The reference's BIF loader needs Boost.Spirit and cannot
be built here: BIF parity is pinned on its grammar (bif.hpp:138-263) through hand-written files, and on a public
third-party file (tests/golden/asia.bif) through published marginals and the reference DSC loader
-- "parity unpinned" for BIF in the task's terms.
"""
import os

import numpy as np
import pytest

from bayesiannetwork_b200 import netfile, synth
from bayesiannetwork_b200._capi import BnbpError

HERE = os.path.dirname(os.path.abspath(__file__))


def same_net(a, b):
    assert np.array_equal(a.card, b.card)
    assert np.array_equal(a.parent_off, b.parent_off)
    assert np.array_equal(a.parents, b.parents)
    assert np.array_equal(a.cpt_off, b.cpt_off)
    assert np.array_equal(a.cpt, b.cpt)          # bit for bit: 17 significant digits written


NETS = [synth.pearl_network, synth.resume_network, synth.alarm37, lambda: synth.grid(4, seed=3),
        lambda: synth.random_dag(40, seed=5)]


@pytest.mark.parametrize("make", NETS)
@pytest.mark.parametrize("fmt", ["bif", "dsc"])
@pytest.mark.parametrize("order", ["ascending", "reversed"])
def test_round_trip(make, fmt, order):
    net = make()
    text = (netfile.dump_bif if fmt == "bif" else netfile.dump_dsc)(net, order=order)
    for f in (fmt, "auto"):
        got = netfile.loads(text, f)
        same_net(got.net, net)
        assert got.node_names == [f"n{i}" for i in range(net.n_nodes)]
        assert got.state_names[0] == [f"s{k}" for k in range(int(net.card[0]))]


def test_load_by_extension(tmp_path):
    net = synth.alarm37()
    for ext, dump in (("bif", netfile.dump_bif), ("dsc", netfile.dump_dsc)):
        p = tmp_path / f"alarm.{ext}"
        p.write_text(dump(net))
        same_net(netfile.load(str(p)).net, net)
    with pytest.raises(BnbpError, match="cannot open"):
        netfile.load(str(tmp_path / "missing.bif"))


# The reference grammar's own shape (bif.hpp:202-248): names, `table` priors, one line per parent
# configuration with the parents' STATE NAMES, parents listed in an order that is not the vertex order.
SPRINKLER = """
network sprinkler {
}
variable Cloudy {
  type discrete [ 2 ] { no, yes };
}
variable Rain {
  type discrete [ 3 ] { none, light, heavy };
}
variable Sprinkler {
  type discrete [ 2 ] { off, on };
}
variable Wet-Grass {
  type discrete [ 2 ] { dry, wet };
}
probability ( Cloudy ) {
  table 0.5, 0.5;
}
probability ( Rain | Cloudy ) {
  (no) 0.8, 0.15, 0.05;
  (yes) 0.2, 0.5, 0.3;
}
probability ( Sprinkler | Cloudy ) {
  (yes) 0.9, 0.1;
  (no) 0.5, 0.5;
}
probability ( Wet-Grass | Sprinkler, Rain ) {
  (off, none) 1.0, 0.0;
  (off, light) 0.3, 0.7;
  (off, heavy) 0.1, 0.9;
  (on, none) 0.2, 0.8;
  (on, light) 0.05, 0.95;
  (on, heavy) 0.01, 0.99;
}
"""


def test_bif_reference_grammar():
    nf = netfile.loads(SPRINKLER, "bif")
    net = nf.net
    assert net.name == "sprinkler"
    assert nf.node_names == ["Cloudy", "Rain", "Sprinkler", "Wet-Grass"]
    assert nf.state_names[1] == ["none", "light", "heavy"]
    assert list(net.card) == [2, 3, 2, 2]
    assert list(net.parent_off) == [0, 0, 1, 2, 4]
    assert list(net.parents) == [0, 0, 1, 2]            # Wet-Grass: ascending index (Rain, Sprinkler)
    cpt = [net.cpt[net.cpt_off[i]:net.cpt_off[i + 1]] for i in range(4)]
    assert list(cpt[0]) == [0.5, 0.5]
    assert list(cpt[1]) == [0.8, 0.15, 0.05, 0.2, 0.5, 0.3]
    assert list(cpt[2]) == [0.5, 0.5, 0.9, 0.1]         # rows re-sorted: (no) first
    # flat layout: first parent (Rain) slowest, Sprinkler fastest
    assert list(cpt[3]) == [1.0, 0.0, 0.2, 0.8, 0.3, 0.7, 0.05, 0.95, 0.1, 0.9, 0.01, 0.99]
    assert nf.node("Rain") == 1 and nf.state(1, "heavy") == 2


# What published files add to that grammar: comments, property lines, quoted names, numbers
# separated by blanks, `default`, and `table` on a node with parents (own state slowest).
DOG = """
// Charniak's dog problem, JavaBayes layout
network "Dog-Problem" { //5 variables
  property "credal-set constant-density-bounded 1.1" ;
}
variable  "light-on" { //2 values
  type discrete[2] {  "true"  "false" };
  property "position = (218, 195)" ;
}
variable  "family-out" {
  type discrete[2] {  "true"  "false" };
}
variable "hear-bark" { type discrete[2] { "true" "false" }; }
probability (  "light-on"  "family-out" ) { /* 2 variable(s) and 4 values */
  table 0.6 0.05 0.4 0.95 ;
}
probability (  "family-out" ) {
  table 0.15 0.85 ;
}
probability ( "hear-bark" | "family-out", "light-on" ) {
  default 0.5 0.5;
  (true, false) 0.7, 0.3;
}
"""


def test_bif_extensions():
    nf = netfile.loads(DOG)
    net = nf.net
    assert net.name == "Dog-Problem"
    assert nf.node_names == ["light-on", "family-out", "hear-bark"]
    assert list(net.parents) == [1, 0, 1]
    cpt = [net.cpt[net.cpt_off[i]:net.cpt_off[i + 1]] for i in range(3)]
    assert list(cpt[0]) == [0.6, 0.4, 0.05, 0.95]       # P(light-on | family-out): rows sum to 1
    assert list(cpt[1]) == [0.15, 0.85]
    # parents of hear-bark in ascending index: (light-on, family-out); the listed row is
    # family-out=true, light-on=false -> q = 1*2 + 0
    assert list(cpt[2]) == [0.5, 0.5, 0.5, 0.5, 0.7, 0.3, 0.5, 0.5]


@pytest.mark.parametrize("text,why", [
    ("variable a { type discrete [ 2 ] { x, y }; } probability ( b ) { table 0.5, 0.5; }", "unknown variable"),
    ("variable a { type discrete [ 2 ] { x, y }; } probability ( a ) { table 0.5; }", "expected 2"),
    ("variable a { type discrete [ 3 ] { x, y }; }", "differs from its list"),
    ("network a { } network b { }", "too many network"),
    ("variable a { type discrete [ 2 ] { x, y }; } variable b { type discrete [ 2 ] { x, y }; } "
     "probability ( a | b ) { (z) 0.5, 0.5; }", "unknown value"),
    ("variable a { type discrete [ 2 ] { x, y }; } variable b { type discrete [ 2 ] { x, y }; } "
     "probability ( a | b ) { } probability ( b | a ) { }", "closes a cycle"),
    ("variable a { type continuous; }", "expected 'discrete'"),
    ("variable a { type discrete [ 2 ] { x, y }; } probability ( a ) { table 0.5, 0.5; ", "line 1"),
])
def test_bif_errors(text, why):
    with pytest.raises(BnbpError, match=why):
        netfile.loads(text, "bif")


# ---- DSC against the reference's own loader -------------------------------------------------------------
DSC_GOLDEN = os.path.join(HERE, "golden", "dsc_fixture.npz")


def test_dsc_golden_fixture():
    """Fixture written by tests/golden/make_dsc_fixture.py from the reference's own DSC loader."""
    fx = np.load(DSC_GOLDEN, allow_pickle=False)
    for key in sorted(k[:-5] for k in fx.files if k.endswith("/text")):
        got = netfile.loads(str(fx[key + "/text"]), "dsc").net
        assert np.array_equal(got.card, fx[key + "/card"]), key
        assert np.array_equal(got.parent_off, fx[key + "/parent_off"]), key
        assert np.array_equal(got.parents, fx[key + "/parents"]), key
        assert np.array_equal(got.cpt_off, fx[key + "/cpt_off"]), key
        assert np.array_equal(got.cpt, fx[key + "/cpt"]), key


@pytest.mark.parametrize("make", NETS)
@pytest.mark.parametrize("order", ["ascending", "reversed"])
def test_dsc_live_reference(make, order, oracle_mod):
    if not oracle_mod.have_reference_dsc():
        pytest.skip("oracle/_ref/libbnref_dsc.so not built (no /root/reference on this machine)")
    net = make()
    text = netfile.dump_dsc(net, order=order)
    card, poff, par, coff, cpt = oracle_mod.reference_dsc_flatten(text)
    got = netfile.loads(text, "dsc").net
    assert np.array_equal(got.card, card) and np.array_equal(got.parent_off, poff)
    assert np.array_equal(got.parents, par) and np.array_equal(got.cpt_off, coff)
    assert np.array_equal(got.cpt, cpt)


def test_dsc_free_layout():
    """What the reference's column-offset reader cannot take: braces on the same line, attributes over
    several lines, state names in condition tuples, `default`."""
    text = '''belief network "two"
    node A { name: "A"; type: discrete[2] =
      { "lo",
        "hi" };
      position: (10, 20);
    }
    node B { type: discrete[2] = {"f", "t"}; }
    probability(A) { 0.25, 0.75; }
    probability(B | A) {
      default: 0.5, 0.5;
      ("hi"): 0.1, 0.9;
    }'''
    nf = netfile.loads(text)
    assert nf.net.name == "two" and nf.node_names == ["A", "B"] and nf.state_names[0] == ["lo", "hi"]
    assert list(nf.net.cpt) == [0.25, 0.75, 0.5, 0.5, 0.1, 0.9]


# ---- BIF pinned on a third-party file ------------------------------------------------------------------------------
# The reference's BIF loader is a Boost.Spirit Qi grammar (serializer/bif.hpp:138-263): Spirit is a template library
# of tens of thousands of lines and cannot be stood in for (unlike Boost.Test / Boost.Algorithm), so the reference
# parser itself cannot be compiled here.  The BIF loader is pinned instead on a PUBLIC third-party file, the ASIA
# network of Lauritzen & Spiegelhalter (1988) as the bnlearn repository distributes it (tests/golden/asia.bif):
#  (1) the exact marginals of what the loader read equal the published ones (brute-force enumeration, no BP);
#  (2) written as DSC, the reference's OWN DSC loader reads back the same arrays (live and as a committed fixture).
ASIA = os.path.join(HERE, "golden", "asia.bif")
ASIA_PUBLISHED = {"asia": 0.01, "tub": 0.0104, "smoke": 0.5, "lung": 0.055, "bronc": 0.45, "either": 0.064828,
                  "xray": 0.11029004, "dysp": 0.4359706}


def _exact_marginals(net):
    import itertools
    marg = [np.zeros(int(r)) for r in net.card]
    for st in itertools.product(*[range(int(r)) for r in net.card]):
        p = 1.0
        for x in range(net.n_nodes):
            q = 0
            for u in net.parents[net.parent_off[x]:net.parent_off[x + 1]]:
                q = q * int(net.card[u]) + st[u]
            p *= net.cpt[net.cpt_off[x] + q * int(net.card[x]) + st[x]]
        for x in range(net.n_nodes):
            marg[x][st[x]] += p
    return marg


def test_bif_third_party_asia_published_marginals():
    nf = netfile.load(ASIA)
    assert nf.node_names == ["asia", "tub", "smoke", "lung", "bronc", "either", "xray", "dysp"]
    assert all(s == ["yes", "no"] for s in nf.state_names)
    net = nf.net
    # `either | lung, tub` lists its parents in the other order than the variables appear: in_vertexes order wins
    assert list(net.parents[net.parent_off[5]:net.parent_off[6]]) == [1, 3]
    for name, m in zip(nf.node_names, _exact_marginals(net)):
        assert abs(m[0] - ASIA_PUBLISHED[name]) < 5e-8, (name, m)
        assert abs(m.sum() - 1.0) < 1e-12


def test_bif_third_party_asia_agrees_with_the_reference_dsc_loader(oracle_mod):
    net = netfile.load(ASIA).net
    text = netfile.dump_dsc(net)
    same_net(netfile.loads(text, "dsc").net, net)                  # own DSC loader: round trip
    fx = np.load(os.path.join(HERE, "golden", "dsc_fixture.npz"), allow_pickle=False)
    assert str(fx["asia_from_bif/text"]) == text                   # the committed reference output is for this text
    assert np.array_equal(fx["asia_from_bif/card"], net.card) and np.array_equal(fx["asia_from_bif/parents"], net.parents)
    assert np.array_equal(fx["asia_from_bif/parent_off"], net.parent_off) and np.array_equal(fx["asia_from_bif/cpt_off"], net.cpt_off)
    assert np.array_equal(fx["asia_from_bif/cpt"], net.cpt)
    if oracle_mod.have_reference_dsc():                            # live, where /root/reference exists
        card, poff, par, coff, cpt = oracle_mod.reference_dsc_flatten(text)
        assert np.array_equal(card, net.card) and np.array_equal(par, net.parents) and np.array_equal(cpt, net.cpt)


@pytest.mark.gpu
def test_bif_third_party_asia_through_the_gpu_path():
    """ASIA has one loop (smoke -> lung -> either -> dysp <- bronc <- smoke): loopy BP is approximate there, but with no
    evidence every message is a plain marginalisation and the beliefs of the loop-free part are exact."""
    from bayesiannetwork_b200.engine import BeliefPropagation
    from bayesiannetwork_b200.flat import EvidenceBatch
    nf = netfile.load(ASIA)
    res = BeliefPropagation(nf.net)(EvidenceBatch.empty(1), 1e-9, max_sweeps=100)
    off = nf.net.belief_off
    for i, name in enumerate(nf.node_names[:7]):                   # everything but dysp sits above the loop's collider
        assert abs(res.marginals[0, off[i]] - ASIA_PUBLISHED[name]) < 1e-7, name
