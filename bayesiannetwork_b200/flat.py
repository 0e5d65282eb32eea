"""Flat (CSR) network and evidence containers shared by the host code, the C ABI and the tests.

The layout is exactly ``bnbp_flat_network`` / ``bnbp_evidence`` of ``include/bnbp.h``:

* node ``i`` is the ``i``-th vertex of the reference's ``graph_t::vertex_list()`` (graph.hpp:214),
* parents are ``graph_t::in_vertexes`` in ascending vertex index (graph.hpp:389-413),
* the CPT of node ``i`` is row-major ``cpt[cpt_off[i] + q*card[i] + x]`` with the parent
  configuration ``q`` in mixed radix, FIRST parent slowest -- the enumeration order of
  ``all_combination_pattern`` (belief_propagation.hpp:269-295).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np


@dataclass
class FlatNetwork:
    card: np.ndarray          # int32 [N]
    parent_off: np.ndarray    # int32 [N+1]
    parents: np.ndarray       # int32 [E]
    cpt_off: np.ndarray       # int64 [N+1]
    cpt: np.ndarray           # float64 [cpt_off[N]]
    name: str = "net"

    def __post_init__(self):
        self.card = np.ascontiguousarray(self.card, dtype=np.int32)
        self.parent_off = np.ascontiguousarray(self.parent_off, dtype=np.int32)
        self.parents = np.ascontiguousarray(self.parents, dtype=np.int32)
        self.cpt_off = np.ascontiguousarray(self.cpt_off, dtype=np.int64)
        self.cpt = np.ascontiguousarray(self.cpt, dtype=np.float64)
        self.validate()

    # ---- derived sizes -------------------------------------------------
    @property
    def n_nodes(self) -> int:
        return int(self.card.shape[0])

    @property
    def n_edges(self) -> int:
        return int(self.parents.shape[0])

    @property
    def belief_off(self) -> np.ndarray:
        out = np.zeros(self.n_nodes + 1, dtype=np.int64)
        np.cumsum(self.card, out=out[1:])
        return out

    @property
    def belief_values(self) -> int:
        """sum_X r_X: width of one row of marginals."""
        return int(self.card.sum())

    @property
    def msg_values(self) -> int:
        """2 * sum_{U->X} r_U: pi- and lambda-message entries of one case."""
        return int(2 * self.card[self.parents].sum()) if self.n_edges else 0

    @property
    def state_values(self) -> int:
        """S = 2*sum r_X + 2*sum_{U->X} r_U."""
        return 2 * self.belief_values + self.msg_values

    def n_configs(self, x: int) -> int:
        q = 1
        for u in self.parents[self.parent_off[x]:self.parent_off[x + 1]]:
            q *= int(self.card[u])
        return q

    def validate(self) -> None:
        n = self.n_nodes
        if self.parent_off.shape[0] != n + 1 or self.cpt_off.shape[0] != n + 1:
            raise ValueError("offset arrays must have n_nodes+1 entries")
        if self.parent_off[0] != 0 or self.cpt_off[0] != 0:
            raise ValueError("offset arrays must start at 0")
        if np.any(self.card < 1):
            raise ValueError("every node needs at least one state")
        if self.parent_off[-1] != self.parents.shape[0]:
            raise ValueError("parent_off[-1] != len(parents)")
        if self.cpt_off[-1] != self.cpt.shape[0]:
            raise ValueError("cpt_off[-1] != len(cpt)")
        for x in range(n):
            ps = self.parents[self.parent_off[x]:self.parent_off[x + 1]]
            if ps.size and (ps.min() < 0 or ps.max() >= n or np.any(ps == x)):
                raise ValueError(f"node {x}: parent id out of range")
            if ps.size > 1 and np.any(np.diff(ps) <= 0):
                raise ValueError(f"node {x}: parents must be strictly ascending (in_vertexes order)")
            q = 1
            for u in ps:
                q *= int(self.card[u])
            if self.cpt_off[x + 1] - self.cpt_off[x] != q * int(self.card[x]):
                raise ValueError(f"node {x}: CPT has {self.cpt_off[x+1]-self.cpt_off[x]} entries, "
                                 f"expected {q}*{int(self.card[x])}")
        # acyclicity (graph_t::add_edge refuses cycles, graph.hpp:268-275)
        indeg = np.diff(self.parent_off).astype(np.int64)
        order = [int(i) for i in np.nonzero(indeg == 0)[0]]
        children = [[] for _ in range(n)]
        for x in range(n):
            for u in self.parents[self.parent_off[x]:self.parent_off[x + 1]]:
                children[int(u)].append(x)
        seen = 0
        while order:
            u = order.pop()
            seen += 1
            for c in children[u]:
                indeg[c] -= 1
                if indeg[c] == 0:
                    order.append(c)
        if seen != n:
            raise ValueError("network is not a DAG")

    # ---- construction helper -------------------------------------------------------------------
    @staticmethod
    def from_lists(card: Sequence[int], parents: Sequence[Sequence[int]],
                   cpts: Sequence[Sequence[float]], name: str = "net") -> "FlatNetwork":
        """cpts[i] is the flattened [Q][r] table of node i (first parent slowest)."""
        n = len(card)
        poff = np.zeros(n + 1, dtype=np.int32)
        coff = np.zeros(n + 1, dtype=np.int64)
        flat_p, flat_c = [], []
        for i in range(n):
            ps = list(parents[i])
            flat_p.extend(ps)
            poff[i + 1] = len(flat_p)
            flat_c.extend(float(v) for v in cpts[i])
            coff[i + 1] = len(flat_c)
        return FlatNetwork(np.asarray(card, np.int32), poff, np.asarray(flat_p, np.int32),
                           coff, np.asarray(flat_c, np.float64), name=name)


@dataclass
class EvidenceBatch:
    """CSR evidence over cases (``bnbp_evidence``).  Hard evidence = one-hot state per entry,
    soft evidence = a full row per entry (the reference's vertex -> 1 x r matrix map)."""
    n_cases: int
    ev_off: np.ndarray                    # int64 [n_cases+1]
    ev_node: np.ndarray                   # int32 [nnz]
    ev_state: Optional[np.ndarray] = None # int32 [nnz]      (hard)
    ev_val_off: Optional[np.ndarray] = None  # int64 [nnz+1] (soft)
    ev_values: Optional[np.ndarray] = None   # float64       (soft)

    def __post_init__(self):
        self.ev_off = np.ascontiguousarray(self.ev_off, dtype=np.int64)
        self.ev_node = np.ascontiguousarray(self.ev_node, dtype=np.int32)
        if self.ev_values is not None:
            self.ev_val_off = np.ascontiguousarray(self.ev_val_off, dtype=np.int64)
            self.ev_values = np.ascontiguousarray(self.ev_values, dtype=np.float64)
            self.ev_state = None
        else:
            if self.ev_state is None:
                self.ev_state = np.zeros(0, dtype=np.int32)
            self.ev_state = np.ascontiguousarray(self.ev_state, dtype=np.int32)
        if self.ev_off.shape[0] != self.n_cases + 1:
            raise ValueError("ev_off must have n_cases+1 entries")

    @property
    def nnz(self) -> int:
        return int(self.ev_node.shape[0])

    @property
    def is_soft(self) -> bool:
        return self.ev_values is not None

    def nbytes(self) -> int:
        b = self.ev_off.nbytes + self.ev_node.nbytes
        if self.is_soft:
            b += self.ev_val_off.nbytes + self.ev_values.nbytes
        else:
            b += self.ev_state.nbytes
        return b

    def slice(self, lo: int, hi: int) -> "EvidenceBatch":
        """Cases [lo, hi) as a self-contained batch (used to shard across GPUs)."""
        a, b = int(self.ev_off[lo]), int(self.ev_off[hi])
        off = self.ev_off[lo:hi + 1] - a
        if self.is_soft:
            va, vb = int(self.ev_val_off[a]), int(self.ev_val_off[b])
            return EvidenceBatch(hi - lo, off, self.ev_node[a:b], None,
                                 self.ev_val_off[a:b + 1] - va, self.ev_values[va:vb])
        return EvidenceBatch(hi - lo, off, self.ev_node[a:b], self.ev_state[a:b])

    @staticmethod
    def empty(n_cases: int) -> "EvidenceBatch":
        return EvidenceBatch(n_cases, np.zeros(n_cases + 1, np.int64), np.zeros(0, np.int32),
                             np.zeros(0, np.int32))

    @staticmethod
    def from_cases(net: FlatNetwork, cases: Sequence[dict]) -> "EvidenceBatch":
        """cases[c] maps node -> state (int, hard) or node -> row (sequence, soft).  If any entry
        is a row, the whole batch is stored as soft evidence (ints become one-hot rows)."""
        soft = any(not isinstance(v, (int, np.integer)) for c in cases for v in c.values())
        off = [0]
        nodes, states, voff, vals = [], [], [0], []
        for c in cases:
            for node, v in c.items():
                r = int(net.card[node])
                nodes.append(int(node))
                if soft:
                    if isinstance(v, (int, np.integer)):
                        row = [1.0 if i == int(v) else 0.0 for i in range(r)]
                    else:
                        row = [float(t) for t in v]
                        if len(row) != r:
                            raise ValueError(f"evidence row for node {node} needs {r} entries")
                    vals.extend(row)
                    voff.append(len(vals))
                else:
                    if not 0 <= int(v) < r:
                        raise ValueError(f"evidence state {v} out of range for node {node}")
                    states.append(int(v))
            off.append(len(nodes))
        if soft:
            return EvidenceBatch(len(cases), np.asarray(off), np.asarray(nodes, np.int32), None,
                                 np.asarray(voff, np.int64), np.asarray(vals, np.float64))
        return EvidenceBatch(len(cases), np.asarray(off), np.asarray(nodes, np.int32),
                             np.asarray(states, np.int32))
