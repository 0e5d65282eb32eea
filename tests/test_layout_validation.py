"""What `bnbp_create` checks before it lays a network out -- the checks the reference leaves as undefined behaviour
(graph.hpp:117-124: a missing CPT row is a dangling reference; `add_edge` is the only place cycles are refused).  The
layout pass is host code, so it is exercised here WITHOUT a GPU through `bnbp_spec_source`, which runs the same
`build_layout()` and then the network compiler; the arrays are handed over raw (the Python container would refuse
most of them itself)."""
import ctypes as C

import numpy as np
import pytest

from bayesiannetwork_b200 import _capi, synth


def _call(card, parent_off, parents, cpt_off, cpt):
    lib = _capi.load()
    arrs = [np.ascontiguousarray(card, np.int32), np.ascontiguousarray(parent_off, np.int32),
            np.ascontiguousarray(parents if len(parents) else [0], np.int32), np.ascontiguousarray(cpt_off, np.int64),
            np.ascontiguousarray(cpt, np.float64)]
    net = _capi.FlatNetworkC(len(card), *[a.ctypes.data_as(C.c_void_p) for a in arrs])
    opt = _capi.OptionsC(0, -1, 0, 0)
    need = C.c_int64(0)
    rc = lib.bnbp_spec_source(C.byref(net), C.byref(opt), 0, None, 0, C.byref(need))
    lib.bnbp_last_error.restype = C.c_char_p
    return rc, lib.bnbp_last_error().decode()


def _chain():
    """A -> B -> C, two states each: a valid network the cases below damage one field at a time."""
    card = [2, 2, 2]
    parent_off = [0, 0, 1, 2]
    parents = [0, 1]
    cpt_off = [0, 2, 6, 10]
    cpt = [0.3, 0.7, 0.9, 0.1, 0.2, 0.8, 0.6, 0.4, 0.5, 0.5]
    return card, parent_off, parents, cpt_off, cpt


def test_valid_network_passes():
    rc, _ = _call(*_chain())
    assert rc == 0


@pytest.mark.parametrize("damage,why", [
    (lambda c, po, p, co, t: (c[:1] + [0] + c[2:], po, p, co, t), "cardinality < 1"),
    (lambda c, po, p, co, t: (c, po, [0, 7], co, t), "bad parent id"),
    (lambda c, po, p, co, t: (c, po, [0, 2], co, t), "bad parent id"),                       # a node as its own parent
    (lambda c, po, p, co, t: (c, [0, 0, 1, 0], p, co, t), "parent_off not monotone"),
    (lambda c, po, p, co, t: (c, po, p, [0, 2, 6, 8], t), "CPT size does not match"),        # a missing CPT row
    (lambda c, po, p, co, t: ([2, 2, 129], po, p, [0, 2, 6, 264], t + [0.0] * 254), "cardinality > 128"),
])
def test_damaged_networks_are_refused_with_the_reason(damage, why):
    rc, msg = _call(*damage(*_chain()))
    assert rc != 0 and why in msg, (rc, msg)


def test_cycles_unsorted_parents_and_wide_parent_sets_are_refused():
    # A <-> B: each lists the other as its parent
    rc, msg = _call([2, 2], [0, 1, 2], [1, 0], [0, 4, 8], [0.5] * 8)
    assert rc != 0 and "cycle" in msg, msg
    # parents must come in ascending vertex index (graph_t::in_vertexes, graph.hpp:389-413)
    rc, msg = _call([2, 2, 2], [0, 0, 0, 2], [1, 0], [0, 2, 4, 12], [0.5] * 12)
    assert rc != 0 and "strictly ascending" in msg, msg
    # more parents than the sweep kernels unroll (documented limit: in-degree <= 8)
    n = 10
    card = [2] * n
    parent_off = [0] * n + [9]
    cpt_off = list(range(0, 2 * n, 2)) + [2 * (n - 1) + 2 * 512]
    rc, msg = _call(card, parent_off, list(range(9)), cpt_off, [0.5] * cpt_off[-1])
    assert rc != 0 and "in-degree > 8" in msg, msg


def test_the_python_container_agrees_on_what_a_valid_network_is():
    net = synth.random_dag(30, 3, 2, 4, seed=2)
    rc, _ = _call(net.card, net.parent_off, net.parents, net.cpt_off, net.cpt)
    assert rc == 0
