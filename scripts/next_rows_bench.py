"""Measurement of the SURVEY 8 'next' rows that are built: likelihood weighting (f2) and CPT estimation (f3).
GPU kernels timed through the C ABI (host buffers in and out), the C restatements (oracle/) timed beside them
on one host core.  Prints one JSON line per row."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bayesiannetwork_b200 import synth
from bayesiannetwork_b200.engine import BeliefPropagation, estimate_cpt
from oracle import oracle

net = synth.alarm37()

# ---- f2: likelihood weighting ------------------------------------------------------------------------
n_cases, n_samples = 4096, 4096
ev = synth.make_evidence(net, n_cases, exact_k=4, seed=5)
bp = BeliefPropagation(net)
bp.likelihood_weighting(ev, 64, seed=1)                       # warm-up (uploads)
t0 = time.perf_counter()
lw = bp.likelihood_weighting(ev, n_samples, seed=2)
dt = time.perf_counter() - t0
small = ev.slice(0, 8)
t0 = time.perf_counter()
ref, _ = oracle.run_port_lw(net, small, n_samples, seed=2)
dt_cpu = time.perf_counter() - t0
assert np.allclose(lw[:8], ref, rtol=1e-10, atol=1e-12)
exact = bp(ev, 1e-9, max_sweeps=200).marginals                # loopy BP on the same evidence: the cross-check itself
print(json.dumps({"row": "f2 likelihood weighting", "workload": f"alarm37, {n_cases} cases x {n_samples} samples",
                  "gpu_node_samples_per_s": n_cases * n_samples * net.n_nodes / dt, "gpu_ms": 1e3 * dt,
                  "cpu_port_node_samples_per_s_1_core": 8 * n_samples * net.n_nodes / dt_cpu,
                  "max_abs_diff_lw_vs_loopy_bp": float(np.abs(lw - exact).max()),
                  "mean_abs_diff_lw_vs_loopy_bp": float(np.abs(lw - exact).mean())}))

# ---- f3: CPT estimation ------------------------------------------------------------------------------
rng = np.random.default_rng(3)
n_rows = 4_000_000
s = np.empty((n_rows, net.n_nodes), dtype=np.int32)
for x in range(net.n_nodes):
    s[:, x] = rng.integers(0, net.card[x], n_rows)
estimate_cpt(net, s[:1000])                                   # warm-up
t0 = time.perf_counter()
got = estimate_cpt(net, s)
dt = time.perf_counter() - t0
t0 = time.perf_counter()
want = oracle.port_make_cpt(net, s[:400_000])
dt_cpu = time.perf_counter() - t0
assert np.array_equal(estimate_cpt(net, s[:400_000]), want)
print(json.dumps({"row": "f3 CPT estimation", "workload": f"alarm37 topology, {n_rows} sample rows x {net.n_nodes} nodes (host table, {s.nbytes / 1e6:.0f} MB)",
                  "gpu_rows_per_s_incl_h2d": n_rows / dt, "gpu_ms": 1e3 * dt, "h2d_floor_ms_at_55GBs": 1e3 * s.nbytes / 55e9,
                  "cpu_port_rows_per_s_1_core": 400_000 / dt_cpu}))
