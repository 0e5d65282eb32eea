"""Fast GPU sanity check (seconds): a few fixtures + a mid-size alarm37 batch vs the oracle.
Run under `timeout` before spending GPU minutes on the full suite."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bayesiannetwork_b200 import synth
from bayesiannetwork_b200.engine import BeliefPropagation
from oracle import oracle

for name, net, kw, eps, cap in [("alarm37", synth.alarm37(), dict(exact_k=4), 1e-6, 200),
                                ("grid8", synth.grid(8, seed=4), dict(p=0.1), 0.0, 30),
                                ("dag80", synth.random_dag(80, 4, 2, 8, seed=7), dict(p=0.1), 0.0, 10)]:
    ev = synth.make_evidence(net, 700, seed=17, **kw)
    om, osw, _ = oracle.run_port(net, ev, eps=eps, max_sweeps=cap, threads=0)
    for prec, tol in (("fp64", 1e-9), ("fp32", 1e-5)):
        if prec == "fp32" and eps > 0:
            continue
        t = time.time()
        r = BeliefPropagation(net, prec)(ev, eps, max_sweeps=cap)
        err = np.nanmax(np.abs(r.marginals - om) / (np.maximum(np.abs(om), np.abs(r.marginals)) + 1e-300 + (1e-12 if prec == "fp64" else 1e-7) / tol))
        ok = np.array_equal(r.sweeps, osw) and err <= tol
        print(f"{name:8s} {prec} sweeps_equal={np.array_equal(r.sweeps, osw)} relerr={err:.2e} {'OK' if ok else 'FAIL'} {time.time()-t:.2f}s", flush=True)
        assert ok
print("sanity ok")
