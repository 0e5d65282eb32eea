"""Network files -> FlatNetwork.

Reading goes through the C ABI (``bnbp_netfile_*`` in ``include/bnbp.h``), i.e. through the drop-in
C++ parsers ``include/bayesian/serializer/{bif,dsc}.hpp`` that replace the reference loaders
(bayesian/serializer/bif.hpp:41-132, dsc.hpp:33-232): one parser, used by C++ hosts and by this
mirror alike.  Host-only code, no GPU needed.

The writers below are new (the reference only reads these formats); they emit text both the
reference grammar and the parsers here accept, with 17 significant digits so CPTs round-trip
bit for bit.  Parents may be listed in any order in a file: ``dump_bif(order="reversed")`` writes
them last-first so tests can check that the loader re-sorts CPT rows into the flat layout.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import _capi
from .flat import FlatNetwork

FORMATS = {"auto": 0, "bif": 1, "dsc": 2}


@dataclass
class NetworkFile:
    net: FlatNetwork
    node_names: List[str]
    state_names: List[List[str]]

    def node(self, name: str) -> int:
        return self.node_names.index(name)

    def state(self, node: int, name: str) -> int:
        return self.state_names[node].index(name)


def _bind(lib):
    if getattr(lib, "_netfile_bound", False):
        return
    lib.bnbp_netfile_parse.restype = C.c_int
    lib.bnbp_netfile_parse.argtypes = [C.c_char_p, C.c_int64, C.c_int32, C.POINTER(C.c_void_p)]
    lib.bnbp_netfile_load.restype = C.c_int
    lib.bnbp_netfile_load.argtypes = [C.c_char_p, C.c_int32, C.POINTER(C.c_void_p)]
    lib.bnbp_netfile_network.restype = C.POINTER(_capi.FlatNetworkC)
    lib.bnbp_netfile_network.argtypes = [C.c_void_p]
    lib.bnbp_netfile_name.restype = C.c_char_p
    lib.bnbp_netfile_name.argtypes = [C.c_void_p]
    lib.bnbp_netfile_node_name.restype = C.c_char_p
    lib.bnbp_netfile_node_name.argtypes = [C.c_void_p, C.c_int32]
    lib.bnbp_netfile_state_name.restype = C.c_char_p
    lib.bnbp_netfile_state_name.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
    lib.bnbp_netfile_free.restype = None
    lib.bnbp_netfile_free.argtypes = [C.c_void_p]
    lib._netfile_bound = True


def _take(lib, handle) -> NetworkFile:
    try:
        v = lib.bnbp_netfile_network(handle).contents
        n = int(v.n_nodes)

        def arr(ptr, count, ctype, dtype):
            if count == 0:
                return np.zeros(0, dtype=dtype)
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(count,)).astype(dtype, copy=True)

        card = arr(v.card, n, C.c_int32, np.int32)
        poff = arr(v.parent_off, n + 1, C.c_int32, np.int32)
        parents = arr(v.parents, int(poff[-1]), C.c_int32, np.int32)
        coff = arr(v.cpt_off, n + 1, C.c_int64, np.int64)
        cpt = arr(v.cpt, int(coff[-1]), C.c_double, np.float64)
        name = lib.bnbp_netfile_name(handle).decode("utf-8", "replace") or "net"
        nodes = [lib.bnbp_netfile_node_name(handle, i).decode("utf-8", "replace") for i in range(n)]
        states = [[lib.bnbp_netfile_state_name(handle, i, s).decode("utf-8", "replace") for s in range(int(card[i]))]
                  for i in range(n)]
        return NetworkFile(FlatNetwork(card, poff, parents, coff, cpt, name=name), nodes, states)
    finally:
        lib.bnbp_netfile_free(handle)


def loads(text: str, fmt: str = "auto") -> NetworkFile:
    """Parse BIF or DSC text (``fmt`` in auto / bif / dsc)."""
    lib = _capi.load()
    _bind(lib)
    raw = text.encode("utf-8")
    handle = C.c_void_p()
    _capi.check(lib.bnbp_netfile_parse(raw, len(raw), FORMATS[fmt], C.byref(handle)))
    return _take(lib, handle)


def load(path: str, fmt: str = "auto") -> NetworkFile:
    lib = _capi.load()
    _bind(lib)
    handle = C.c_void_p()
    _capi.check(lib.bnbp_netfile_load(path.encode("utf-8"), FORMATS[fmt], C.byref(handle)))
    return _take(lib, handle)


# ---- writers ------------------------------------------------------------------------------------------
def _names(net: FlatNetwork, node_names, state_names):
    nodes = list(node_names) if node_names is not None else [f"n{i}" for i in range(net.n_nodes)]
    states = ([list(s) for s in state_names] if state_names is not None
              else [[f"s{k}" for k in range(int(net.card[i]))] for i in range(net.n_nodes)])
    return nodes, states


def _rows(net: FlatNetwork, x: int, listed: List[int]):
    """(parent states in `listed` order, row) for every configuration, `listed` order first-slowest."""
    ps = [int(u) for u in net.parents[net.parent_off[x]:net.parent_off[x + 1]]]
    r = int(net.card[x])
    base = int(net.cpt_off[x])
    radix = [int(net.card[u]) for u in listed]
    q_total = int(np.prod(radix)) if radix else 1
    for ql in range(q_total):
        digits, rem = [], ql
        for rj in reversed(radix):
            digits.append(rem % rj)
            rem //= rj
        digits.reverse()
        by_parent = dict(zip(listed, digits))
        q = 0
        for u in ps:
            q = q * int(net.card[u]) + by_parent[u]
        yield digits, net.cpt[base + q * r: base + (q + 1) * r]


def dump_bif(net: FlatNetwork, node_names=None, state_names=None, order: str = "ascending") -> str:
    nodes, states = _names(net, node_names, state_names)
    out = [f"network {net.name or 'unknown'} {{\n}}\n"]
    for i in range(net.n_nodes):
        out.append(f"variable {nodes[i]} {{\n  type discrete [ {int(net.card[i])} ] {{ {', '.join(states[i])} }};\n}}\n")
    for x in range(net.n_nodes):
        ps = [int(u) for u in net.parents[net.parent_off[x]:net.parent_off[x + 1]]]
        listed = ps[::-1] if order == "reversed" else ps
        head = nodes[x] + (" | " + ", ".join(nodes[u] for u in listed) if listed else "")
        out.append(f"probability ( {head} ) {{\n")
        for digits, row in _rows(net, x, listed):
            vals = ", ".join(repr(float(v)) for v in row)
            if listed:
                out.append(f"  ({', '.join(states[u][d] for u, d in zip(listed, digits))}) {vals};\n")
            else:
                out.append(f"  table {vals};\n")
        out.append("}\n")
    return "".join(out)


def dump_dsc(net: FlatNetwork, node_names=None, state_names=None, order: str = "ascending") -> str:
    nodes, states = _names(net, node_names, state_names)
    out = [f'belief network "{net.name or "unknown"}"\n']
    for i in range(net.n_nodes):
        quoted = ", ".join(f'"{s}"' for s in states[i])
        out.append(f'node {nodes[i]}\n{{\n  name: "{nodes[i]}";\n  type: discrete[{int(net.card[i])}] = {{{quoted}}};\n}}\n')
    for x in range(net.n_nodes):
        ps = [int(u) for u in net.parents[net.parent_off[x]:net.parent_off[x + 1]]]
        listed = ps[::-1] if order == "reversed" else ps
        head = nodes[x] + (" | " + ", ".join(nodes[u] for u in listed) if listed else "")
        out.append(f"probability({head})\n{{\n")
        for digits, row in _rows(net, x, listed):
            vals = ", ".join(repr(float(v)) for v in row)
            if listed:
                out.append(f"  ({', '.join(str(d) for d in digits)}): {vals};\n")
            else:
                out.append(f"  {vals};\n")
        out.append("}\n")
    return "".join(out)
