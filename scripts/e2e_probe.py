"""Per-call wall time of the host-buffer call (bnbp_run_batch) over pinned buffers: does it drift with the call count?"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bayesiannetwork_b200 import synth
from bayesiannetwork_b200.engine import BeliefPropagation
from bayesiannetwork_b200.flat import EvidenceBatch

prec = sys.argv[1] if len(sys.argv) > 1 else "fp64"
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 16
net = synth.alarm37()
n = 1 << 20
ev = synth.make_evidence(net, n, exact_k=4)
bp = BeliefPropagation(net, prec)

def pin(a):
    t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype, pin_memory=True)
    t.numpy()[...] = a
    return t

p_off, p_node, p_state = pin(ev.ev_off), pin(ev.ev_node), pin(ev.ev_state)
p_out = torch.empty((n, net.belief_values), dtype=torch.float64, pin_memory=True)
evp = EvidenceBatch(n, p_off.numpy(), p_node.numpy(), p_state.numpy())
out = p_out.numpy()
reuse = len(sys.argv) > 4 and sys.argv[4] == "reuse"
sw = np.empty(n, dtype=np.int32) if reuse else None
cv = np.empty(n, dtype=np.uint8) if reuse else None
import gc
if len(sys.argv) > 3 and sys.argv[3] == "nogc":
    gc.collect()
    gc.disable()
ts = []
for i in range(calls):
    t0 = time.perf_counter()
    bp(evp, 0.0, max_sweeps=20, out=out, out_sweeps=sw, out_converged=cv)
    ts.append(1e3 * (time.perf_counter() - t0))
print("reuse" if reuse else "alloc", "gc", "off" if not gc.isenabled() else "on", "| max %.2f median %.2f |" % (max(ts[2:]), sorted(ts[2:])[len(ts[2:]) // 2]), end=" ")
print(prec, "ms per call:", " ".join(f"{t:.2f}" for t in ts), "| device total of the last call", bp.stats()["last_total_ms"])
