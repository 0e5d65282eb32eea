"""Multi-GPU plumbing on CPU: world_size-2 gloo processes exercise everything of the N>1 path that
is not a kernel -- contiguous case sharding, per-rank evidence generation (identical for any
world size), the convergence-summary all-reduce and the marginal gather order.
The kernels themselves are stood in for by a deterministic function of the evidence so that the
test needs no GPU; the real thing runs under torchrun in bench.py --gpus N."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bayesiannetwork_b200 import dist as bdist
from bayesiannetwork_b200 import synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_marginals(net, ev):
    """Stand-in for the GPU result: row c depends only on case c's evidence entries."""
    V = net.belief_values
    out = np.zeros((ev.n_cases, V))
    off = net.belief_off
    rows = np.repeat(np.arange(ev.n_cases), np.diff(ev.ev_off))
    np.add.at(out, (rows, off[ev.ev_node] + ev.ev_state), 1.0)
    return out


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        net = synth.alarm37()
        lo, hi = bdist.shard_range(n_total, rank, world)
        # each rank generates ITS shard only, from the global case index
        ev = synth.make_evidence(net, hi - lo, exact_k=4, case_offset=lo)
        local = torch.from_numpy(_fake_marginals(net, ev))
        sweeps = torch.full((hi - lo,), 20, dtype=torch.int32)
        conv = torch.ones(hi - lo, dtype=torch.uint8)
        if rank == 1:
            conv[0] = 0                                       # one non-converged case on rank 1
        summary = bdist.reduce_summary(sweeps, conv, torch.zeros(2, dtype=torch.int64))
        full = bdist.gather_marginals(local, n_total)
        if rank == 0:
            q.put((summary.tolist(), full.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [1000, 1001])
def test_two_rank_sharding_summary_and_gather(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    summary, full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert summary == [20 * n_total, 1]
    net = synth.alarm37()
    whole = synth.make_evidence(net, n_total, exact_k=4)      # the unsharded batch
    assert full.shape == (n_total, net.belief_values)
    assert np.array_equal(full, _fake_marginals(net, whole))   # gather order == global case order


def test_shard_ranges_partition_the_batch():
    for n in (0, 1, 7, 1000, 1 << 20):
        for world in (1, 2, 3, 8):
            spans = [bdist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_evidence_is_independent_of_the_shard_it_is_generated_in():
    net = synth.alarm37()
    whole = synth.make_evidence(net, 500, exact_k=4)
    part = synth.make_evidence(net, 200, exact_k=4, case_offset=300)
    ref = whole.slice(300, 500)
    assert np.array_equal(part.ev_node, ref.ev_node) and np.array_equal(part.ev_state, ref.ev_state)
    g = synth.grid(6)
    whole = synth.make_evidence(g, 300, p=0.1)
    part = synth.make_evidence(g, 100, p=0.1, case_offset=200)
    ref = whole.slice(200, 300)
    assert np.array_equal(part.ev_off, ref.ev_off) and np.array_equal(part.ev_node, ref.ev_node)
    assert np.array_equal(part.ev_state, ref.ev_state)
