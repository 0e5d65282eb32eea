"""Device time per sweep of the specialised kernel variants (alarm37, 1M cases): plain / freeze / freeze+check / damping."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesiannetwork_b200 import synth
from bayesiannetwork_b200.engine import BeliefPropagation

prec = sys.argv[1] if len(sys.argv) > 1 else "fp64"
net = synth.alarm37()
n = 1 << 20
ev = synth.make_evidence(net, n, exact_k=4)
bp = BeliefPropagation(net, prec, specialize="always")
dev = torch.device("cuda", 0)
d_off, d_node, d_state = (torch.from_numpy(a).to(dev) for a in (ev.ev_off, ev.ev_node, ev.ev_state))
d_out = torch.empty((n, net.belief_values), dtype=torch.float64 if prec == "fp64" else torch.float32, device=dev)
os.environ["BNBP_NO_COMPACT"] = "1"
for name, kw in [("plain", dict(epsilon=0.0)), ("freeze only (check_interval 1000)", dict(epsilon=1e-300, check_interval=1000)),
                 ("freeze+check", dict(epsilon=1e-300)), ("damping 0.1 (check variant, no test)", dict(epsilon=0.0, damping=0.1))]:
    for _ in range(2):
        bp.run_device(n, d_off, d_node, d_state, d_out, max_sweeps=10, **kw)
        torch.cuda.synchronize()
    st = bp.stats()
    print(f"{prec} {name:40s} {st['last_sweep_ms'] / st['last_sweep_launches']:.3f} ms per sweep ({st['last_sweep_launches']} sweeps, total {st['last_total_ms']:.2f} ms)")
