"""Several GPUs behind one handle (bnbp_create_multi) and what leaves the device (query nodes, float marginals)
through the Python host mirror -- needs a B200 (-m gpu); with one GPU the group has one member and the same code runs.
The C++ side of the same features: tests/cpp/test_multi_gpu.cpp."""
import numpy as np
import pytest

from bayesiannetwork_b200 import synth
from helpers import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def BP():
    from bayesiannetwork_b200.engine import BeliefPropagation
    return BeliefPropagation


@pytest.mark.parametrize("mode", ["fixed", "eps"])
def test_group_handle_equals_one_device_bit_for_bit(BP, mode):
    from bayesiannetwork_b200.engine import device_count
    net = synth.alarm37()
    ev = synth.make_evidence(net, 20011, exact_k=4, seed=5)
    eps, cap = (0.0, 15) if mode == "fixed" else (1e-6, 200)
    one = BP(net)(ev, eps, max_sweeps=cap)
    grp = BP(net, devices=[])                       # every visible device
    res = grp(ev, eps, max_sweeps=cap)
    assert np.array_equal(res.marginals, one.marginals)
    assert np.array_equal(res.sweeps, one.sweeps) and np.array_equal(res.converged, one.converged)
    sm = grp.summary()                              # all-reduced over NCCL inside libbnbp
    assert sm["n_cases"] == ev.n_cases and sm["case_sweeps"] == int(one.sweeps.sum())
    assert sm["not_converged"] == int((one.converged == 0).sum()) and sm["max_sweeps"] == int(one.sweeps.max())
    assert grp.stats()["last_case_sweeps"] == sm["case_sweeps"]
    if device_count() >= 2:
        small = ev.slice(0, 37)                                                  # fewer cases than devices x 32
        two = BP(net, devices=[1, 0])(small, eps, max_sweeps=cap)
        assert np.array_equal(two.marginals, BP(net)(small, eps, max_sweeps=cap).marginals)   # (same kernel family: 37 cases)


def test_query_nodes_and_float_marginals(BP, oracle_mod):
    from bayesiannetwork_b200.engine import BnbpError
    net = synth.alarm37()
    off = net.belief_off
    q = [17, 3, 36, 0]
    cols = np.concatenate([np.arange(off[x], off[x + 1]) for x in q])
    for n, kw in ((5000, {}), (300, {}), (5000, dict(specialize="always")), (5000, dict(devices=[]))):
        ev = synth.make_evidence(net, n, exact_k=4, seed=7)
        bp = BP(net, **kw)
        full = bp(ev, 0.0, max_sweeps=10)
        part = bp(ev, 0.0, max_sweeps=10, query_nodes=q)
        assert part.marginals.shape == (n, len(cols))
        assert np.array_equal(part.marginals, full.marginals[:, cols]), (n, kw)
        again = bp(ev, 0.0, max_sweeps=10)          # back to every node
        assert np.array_equal(again.marginals, full.marginals)
    ev = synth.make_evidence(net, 5000, exact_k=4, seed=7)
    with pytest.raises(BnbpError):
        BP(net)(ev, 0.0, max_sweeps=3, query_nodes=[3, 3])
    with pytest.raises(BnbpError):
        BP(net)(ev, 0.0, max_sweeps=3, query_nodes=[99])
    # float marginals: fp32 handles only; equal to the double marginals of the same handle rounded to float
    b32 = BP(net, "fp32")
    wide = b32(ev, 0.0, max_sweeps=10).marginals
    narrow = b32(ev, 0.0, max_sweeps=10, out_dtype=np.float32).marginals
    assert narrow.dtype == np.float32 and np.array_equal(narrow, wide.astype(np.float32))
    om, _, _ = oracle_mod.run_port(net, ev.slice(0, 256), eps=0.0, max_sweeps=10, threads=0)
    assert_close(narrow[:256], om, rtol=1e-5, atol=1e-7, what="float marginals")
    with pytest.raises(BnbpError):
        BP(net, "fp64")(ev, 0.0, max_sweeps=3, out_dtype=np.float32)
