"""Regenerate tests/golden/sampler_fixture.npz from the REFERENCE ITSELF: bayesian/sampler.hpp compiled in place
(oracle/_ref/libbnref_sampler.so, oracle/Makefile `ref`; Boost replaced by the stand-ins of tests/cpp/boost).
Run in the authoring container only (the GPU box has no /root/reference):
    python tests/golden/make_sampler_fixture.py
Each fixture = a network topology, a table of distinct samples with multiplicities, and the CPT arena the reference's
sampler::make_cpt (sampler.hpp:81-163) computes from it -- through load_sample(table) and through its own sample-file
reader (sampler.hpp:42-76), which must agree."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bayesiannetwork_b200 import synth  # noqa: E402
from oracle import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def cases():
    rng = np.random.default_rng(20261018)
    for name, net, rows in (("pearl", synth.pearl_network(), 40), ("alarm37", synth.alarm37(), 3000),
                            ("dag25_card5", synth.random_dag(25, 3, 2, 5, seed=5), 1500)):
        # few rows for the number of configurations: some parent configurations are never seen (uniform rows)
        s = np.stack([rng.integers(0, int(r), size=rows) for r in net.card], axis=1).astype(np.int32)
        s = np.unique(s, axis=0)
        mult = rng.integers(1, 9, size=s.shape[0]).astype(np.int64)
        yield name, net, s, mult


def main():
    out = {}
    for name, net, s, mult in cases():
        cpt = oracle.reference_make_cpt(net, s, mult)
        with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
            for row, m in zip(s, mult):
                f.write(f"{int(m)} " + "  ".join(str(int(v)) for v in row) + "\n")     # two blanks: token_compress_on
            path = f.name
        cpt_file, size = oracle.reference_make_cpt_from_file(net, path)
        os.unlink(path)
        assert size == int(mult.sum()) and np.array_equal(cpt, cpt_file), name
        p = name + "/"
        out.update({p + "card": net.card, p + "parent_off": net.parent_off, p + "parents": net.parents, p + "cpt_off": net.cpt_off,
                    p + "samples": s, p + "mult": mult, p + "cpt": cpt})
        print(name, s.shape, "unseen rows:", int((np.abs(cpt.reshape(-1)[:0]) > 0).sum()), "ok")
    np.savez_compressed(os.path.join(HERE, "sampler_fixture.npz"), **out)


if __name__ == "__main__":
    main()
