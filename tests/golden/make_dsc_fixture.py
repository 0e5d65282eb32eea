"""Writes tests/golden/dsc_fixture.npz: DSC texts and the flat arrays the REFERENCE's own DSC loader
(bayesian/serializer/dsc.hpp, compiled in place as oracle/_ref/libbnref_dsc.so) makes of them.
Run where /root/reference exists:  python tests/golden/make_dsc_fixture.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from bayesiannetwork_b200 import netfile, synth  # noqa: E402
from oracle import oracle  # noqa: E402

if __name__ == "__main__":
    oracle.build()
    assert oracle.have_reference_dsc(), "needs /root/reference"
    cases = {
        "pearl": netfile.dump_dsc(synth.pearl_network()),
        "alarm37_reversed": netfile.dump_dsc(synth.alarm37(), order="reversed"),
        "dag30": netfile.dump_dsc(synth.random_dag(30, seed=11)),
        # the public ASIA network (tests/golden/asia.bif, a third-party BIF file), read by the BIF loader, written as
        # DSC: the reference's own DSC loader must give back what the BIF loader read (pins the BIF loader by proxy)
        "asia_from_bif": netfile.dump_dsc(netfile.load(os.path.join(ROOT, "tests", "golden", "asia.bif")).net),
    }
    out = {}
    for name, text in cases.items():
        card, poff, par, coff, cpt = oracle.reference_dsc_flatten(text)
        out[name + "/text"] = np.array(text)
        out[name + "/card"], out[name + "/parent_off"], out[name + "/parents"] = card, poff, par
        out[name + "/cpt_off"], out[name + "/cpt"] = coff, cpt
    path = os.path.join(ROOT, "tests", "golden", "dsc_fixture.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if not k.endswith("text")})
