"""Shared helpers for the parity tests."""
from __future__ import annotations

import os

import numpy as np

from bayesiannetwork_b200.flat import EvidenceBatch, FlatNetwork


def load_fixture(fx, name):
    p = name + "/"
    net = FlatNetwork(fx[p + "card"], fx[p + "parent_off"], fx[p + "parents"], fx[p + "cpt_off"],
                      fx[p + "cpt"], name=name)
    n_cases = fx[p + "ev_off"].shape[0] - 1
    if (p + "ev_values") in fx.files:
        ev = EvidenceBatch(n_cases, fx[p + "ev_off"], fx[p + "ev_node"], None,
                           fx[p + "ev_val_off"], fx[p + "ev_values"])
    else:
        ev = EvidenceBatch(n_cases, fx[p + "ev_off"], fx[p + "ev_node"], fx[p + "ev_state"])
    return dict(net=net, ev=ev, eps=float(fx[p + "eps"]), max_sweeps=int(fx[p + "max_sweeps"]),
                marginals=fx[p + "marginals"], sweeps=fx[p + "sweeps"], converged=fx[p + "converged"])


def assert_close(a, b, rtol, atol, what=""):
    """|a-b| <= rtol*max(|a|,|b|) + atol elementwise, NaN must match NaN (SURVEY 8d parity check)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    na, nb = np.isnan(a), np.isnan(b)
    assert np.array_equal(na, nb), f"{what}: NaN pattern differs ({int(na.sum())} vs {int(nb.sum())})"
    ok = ~na
    err = np.abs(a[ok] - b[ok])
    bound = rtol * np.maximum(np.abs(a[ok]), np.abs(b[ok])) + atol
    bad = err > bound
    log = os.environ.get("BNBP_MARGIN_LOG")          # how close to the bound: one line per comparison
    if log and err.size:
        with open(log, "a") as f:
            f.write(f"{what}: max err/bound {float((err / bound).max()):.3f}, max err {float(err.max()):.3e} "
                    f"(rtol {rtol:g}, atol {atol:g}, {err.size} entries)\n")
    assert not bad.any(), f"{what}: {int(bad.sum())} entries off, max err {err.max():.3e}"
