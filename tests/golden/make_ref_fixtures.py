"""Regenerate tests/golden/ref_fixtures.npz from the REFERENCE ITSELF (oracle/_ref/libbnref.so,
the unmodified headers under /root/reference compiled in place by oracle/Makefile).

Run in the authoring container only (the GPU box has no /root/reference):
    python tests/golden/make_ref_fixtures.py            # writes ref_fixtures.npz
    python tests/golden/make_ref_fixtures.py --check-json   # also re-derives reference_tests.json numbers

Each fixture = a flat network, an evidence batch, (eps, max_sweeps) and the reference's marginals,
sweep counts and converged flags.  The fixtures deliberately cover what the reference's own tests
never do: loopy graphs, soft evidence, impossible evidence (NaN), fixed sweep counts.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from bayesiannetwork_b200 import synth  # noqa: E402
from bayesiannetwork_b200.flat import EvidenceBatch  # noqa: E402
from oracle import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def fixtures():
    pearl, resume = synth.pearl_network(), synth.resume_network()
    yield "pearl_tests", pearl, EvidenceBatch.from_cases(pearl, [{}, {3: 0}]), 1e-3, 0
    yield "resume_tests", resume, EvidenceBatch.from_cases(
        resume, [{1: 2, 3: 0}, {2: 1}, {0: 1, 2: 1}, {3: 2}, {0: 0}]), 1e-3, 0
    yield "resume_soft", resume, EvidenceBatch.from_cases(resume, [{2: [0.3, 0.7]}, {1: [0.2, 0.5, 0.3], 3: [0.1, 0.1, 0.8]}]), 1e-3, 0
    # impossible evidence: Pearl with R=0 and W=1 (P(W=1|R=0)=0) -> NaN messages
    yield "pearl_nan", pearl, EvidenceBatch.from_cases(pearl, [{0: 0, 2: 1}, {0: 0, 2: 1, 3: 1}]), 1e-3, 50
    yield "pearl_nan_fixed6", pearl, EvidenceBatch.from_cases(pearl, [{0: 0, 2: 1}, {2: 1}, {0: 0, 2: 1, 3: 1}]), 0.0, 6
    yield "pearl_one_sweep", pearl, EvidenceBatch.from_cases(pearl, [{3: 0}, {}]), 1e300, 0
    poly = synth.random_polytree(24, card_hi=4, seed=11)
    yield "polytree24_eps", poly, synth.make_evidence(poly, 12, p=0.15, seed=5), 1e-9, 200
    yield "polytree24_soft", poly, synth.make_evidence(poly, 6, p=0.2, seed=6, soft=True), 1e-9, 200
    g4 = synth.grid(4, seed=99)
    yield "grid4_eps", g4, synth.make_evidence(g4, 10, p=0.2, seed=7), 1e-9, 500
    yield "grid4_fixed7", g4, synth.make_evidence(g4, 6, p=0.2, seed=8), 0.0, 7
    g6 = synth.grid(6, seed=100)
    yield "grid6_fixed12", g6, synth.make_evidence(g6, 4, p=0.1, seed=9), 0.0, 12
    dag = synth.random_dag(40, max_parents=4, card_lo=2, card_hi=5, seed=123)
    yield "dag40_eps", dag, synth.make_evidence(dag, 8, p=0.1, seed=10), 1e-6, 300
    yield "dag40_fixed5", dag, synth.make_evidence(dag, 6, p=0.15, seed=11), 0.0, 5
    a = synth.alarm37()
    yield "alarm37_eps", a, synth.make_evidence(a, 16, exact_k=4, seed=1), 1e-6, 200
    yield "alarm37_fixed20", a, synth.make_evidence(a, 8, exact_k=4, seed=1, case_offset=16), 0.0, 20
    hc = synth.high_card(8, card=6, n_parents=3, seed=77)
    yield "card6_fixed6", hc, synth.make_evidence(hc, 4, p=0.2, seed=12), 0.0, 6


def main():
    if not oracle.have_reference():
        raise SystemExit("oracle/_ref/libbnref.so missing: run `make -C oracle` where /root/reference exists")
    out = {}
    names = []
    for name, net, ev, eps, cap in fixtures():
        m, s, c, _ = oracle.run_reference(net, ev, eps=eps, max_sweeps=cap)
        names.append(name)
        p = name + "/"
        out[p + "card"], out[p + "parent_off"], out[p + "parents"] = net.card, net.parent_off, net.parents
        out[p + "cpt_off"], out[p + "cpt"] = net.cpt_off, net.cpt
        out[p + "ev_off"], out[p + "ev_node"] = ev.ev_off, ev.ev_node
        if ev.is_soft:
            out[p + "ev_val_off"], out[p + "ev_values"] = ev.ev_val_off, ev.ev_values
        else:
            out[p + "ev_state"] = ev.ev_state
        out[p + "eps"], out[p + "max_sweeps"] = np.float64(eps), np.int32(cap)
        out[p + "marginals"], out[p + "sweeps"], out[p + "converged"] = m, s, c
        print(f"{name:20s} N={net.n_nodes:3d} cases={ev.n_cases:3d} sweeps={s.tolist()} nan={int(np.isnan(m).sum())}")
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "ref_fixtures.npz"), **out)
    if "--check-json" in sys.argv:
        spec = json.load(open(os.path.join(HERE, "reference_tests.json")))
        nets = {"pearl": synth.pearl_network(), "resume": synth.resume_network()}
        for case in spec["cases"]:
            net = nets[case["network"]]
            ev = EvidenceBatch.from_cases(net, [{int(k): v for k, v in case["evidence"].items()}])
            m, s, _, _ = oracle.run_reference(net, ev, eps=case["eps"], pure=True)
            m2, s2, _, _ = oracle.run_reference(net, ev, eps=case["eps"])
            assert np.allclose(m, m2, rtol=1e-13, atol=1e-16) and s2[0] == case["sweeps"], case["name"]
            if case["beliefs"] is not None:
                flat = np.concatenate([np.asarray(b) for b in case["beliefs"]])
                assert np.allclose(m[0], flat, rtol=0, atol=2e-16), (case["name"], m[0], flat)
            print("json ok:", case["name"], s2[0])


if __name__ == "__main__":
    main()
