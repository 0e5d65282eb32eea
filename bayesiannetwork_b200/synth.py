"""Synthetic networks and evidence of the shapes BASELINE.json names.

Everything is driven by splitmix64 with explicit arithmetic (no library distributions), so the
same seed gives the same network / evidence on every machine, and evidence for case ``c`` depends
only on ``(seed, c)`` -- a shard of cases is identical whichever GPU generates it.

Also holds the two networks of the reference's own BP tests (cfg 1): the Pearl 4-node network
and the "resume" chain (libs/bayesian/test/belief_propagation.cpp:9-62, :127-182).
"""
from __future__ import annotations

import numpy as np

from .flat import EvidenceBatch, FlatNetwork

_M64 = (1 << 64) - 1
_GAMMA = 0x9E3779B97F4A7C15
NETWORK_SEED = 20261017
EVIDENCE_SEED = 1


def _mix_int(z: int) -> int:
    z &= _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


class SplitMix64:
    """Scalar splitmix64 stream (network structure)."""

    def __init__(self, seed: int):
        self.s = seed & _M64

    def next_u64(self) -> int:
        self.s = (self.s + _GAMMA) & _M64
        return _mix_int(self.s)

    def below(self, n: int) -> int:
        return self.next_u64() % n

    def unit(self) -> float:
        return (self.next_u64() >> 11) * (1.0 / 9007199254740992.0)


def _mix_np(z: np.ndarray) -> np.ndarray:
    z = z.astype(np.uint64, copy=True)
    z ^= z >> np.uint64(30)
    z *= np.uint64(0xBF58476D1CE4E5B9)
    z ^= z >> np.uint64(27)
    z *= np.uint64(0x94D049BB133111EB)
    z ^= z >> np.uint64(31)
    return z


def counter_u64(seed: int, index: np.ndarray, draw: int = 0) -> np.ndarray:
    """Counter-mode splitmix64: the ``draw``-th output of the stream seeded with
    ``mix(seed ^ index)`` -- vectorised over ``index``."""
    with np.errstate(over="ignore"):
        s0 = _mix_np(np.uint64(seed & _M64) ^ index.astype(np.uint64))
        return _mix_np(s0 + np.uint64(((draw + 1) * _GAMMA) & _M64))


def _cpt_rows(seed: int, n_rows: int, r: int) -> np.ndarray:
    """n_rows x r table, entries 0.05 + 0.95*u, each row divided by its sum (strictly positive:
    no impossible evidence, loopy grids converge -- SURVEY.md section 6)."""
    idx = np.arange(n_rows * r, dtype=np.uint64)
    u = (counter_u64(seed, idx) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    t = (0.05 + 0.95 * u).reshape(n_rows, r)
    return t / t.sum(axis=1, keepdims=True)


def _assemble(card, parent_lists, seed: int, name: str) -> FlatNetwork:
    n = len(card)
    card = np.asarray(card, dtype=np.int32)
    poff = np.zeros(n + 1, dtype=np.int32)
    coff = np.zeros(n + 1, dtype=np.int64)
    flat_p = []
    for i in range(n):
        ps = sorted(int(p) for p in parent_lists[i])
        flat_p.extend(ps)
        poff[i + 1] = len(flat_p)
        q = 1
        for p in ps:
            q *= int(card[p])
        coff[i + 1] = coff[i] + q * int(card[i])
    cpt = np.empty(int(coff[-1]), dtype=np.float64)
    for i in range(n):
        r = int(card[i])
        rows = int(coff[i + 1] - coff[i]) // r
        cpt[coff[i]:coff[i + 1]] = _cpt_rows(_mix_int(seed ^ (0xC0FFEE + i)), rows, r).ravel()
    return FlatNetwork(card, poff, np.asarray(flat_p, dtype=np.int32), coff, cpt, name=name)


# ---- cfg 1: the reference's own test networks ------------------------------------------------
def pearl_network() -> FlatNetwork:
    """R, S -> W (R), H (R, S); libs/bayesian/test/belief_propagation.cpp:9-62."""
    return FlatNetwork.from_lists(
        card=[2, 2, 2, 2],
        parents=[[], [], [0], [0, 1]],
        cpts=[[0.2, 0.8], [0.1, 0.9],
              [1.0, 0.0, 0.2, 0.8],
              [1.0, 0.0, 1.0, 0.0, 0.9, 0.1, 0.0, 1.0]],
        name="pearl")


def resume_network() -> FlatNetwork:
    """Chain A -> B -> C -> D, cards 3,3,2,3; libs/bayesian/test/belief_propagation.cpp:127-182."""
    return FlatNetwork.from_lists(
        card=[3, 3, 2, 3],
        parents=[[], [0], [1], [2]],
        cpts=[[0.30, 0.60, 0.10],
              [0.20, 0.30, 0.50, 0.30, 0.30, 0.40, 0.80, 0.10, 0.10],
              [0.50, 0.50, 0.70, 0.30, 0.40, 0.60],
              [0.40, 0.30, 0.30, 0.20, 0.60, 0.20]],
        name="resume")


# ---- cfg 2: ALARM-sized --------------------------------------------------------------------------
def alarm37(seed: int = NETWORK_SEED) -> FlatNetwork:
    """37 nodes, exactly 46 edges, card in {2,3,4}, in-degree <= 4: pairs a<b drawn uniformly,
    rejected on duplicate / over-degree."""
    rng = SplitMix64(seed)
    n, n_edges = 37, 46
    card = [2 + rng.below(3) for _ in range(n)]
    parents = [set() for _ in range(n)]
    m = 0
    while m < n_edges:
        a, b = rng.below(n), rng.below(n)
        if a == b:
            continue
        if a > b:
            a, b = b, a
        if a in parents[b] or len(parents[b]) >= 4:
            continue
        parents[b].add(a)
        m += 1
    return _assemble(card, parents, seed, "alarm37")


# ---- cfg 3: loopy grid ---------------------------------------------------------------------------
def grid(n: int = 100, seed: int = NETWORK_SEED) -> FlatNetwork:
    """n x n binary grid, edges (i,j)->(i+1,j) and (i,j)->(i,j+1)."""
    card = [2] * (n * n)
    parents = []
    for i in range(n):
        for j in range(n):
            ps = []
            if i > 0:
                ps.append((i - 1) * n + j)
            if j > 0:
                ps.append(i * n + j - 1)
            parents.append(ps)
    return _assemble(card, parents, seed, f"grid{n}")


# ---- cfg 4: random DAG ---------------------------------------------------------------------------
def random_dag(n: int = 2000, max_parents: int = 4, card_lo: int = 2, card_hi: int = 8,
               seed: int = NETWORK_SEED) -> FlatNetwork:
    """Node i gets min(i, U{0..max_parents}) parents drawn without replacement from its
    predecessors; card uniform in [card_lo, card_hi]."""
    rng = SplitMix64(seed ^ 0xDA6)
    card = [card_lo + rng.below(card_hi - card_lo + 1) for _ in range(n)]
    parents = []
    for i in range(n):
        k = min(i, rng.below(max_parents + 1))
        ps = set()
        while len(ps) < k:
            ps.add(rng.below(i))
        parents.append(ps)
    return _assemble(card, parents, seed, f"dag{n}")


# ---- cfg 5: high cardinality ---------------------------------------------------------------------
def high_card(n: int = 64, card: int = 32, n_parents: int = 3, seed: int = NETWORK_SEED) -> FlatNetwork:
    """Nodes 0..n_parents-1 are roots, every other node has exactly n_parents parents drawn from
    its predecessors."""
    rng = SplitMix64(seed ^ 0xCA4D)
    cards = [card] * n
    parents = []
    for i in range(n):
        ps = set()
        if i >= n_parents:
            while len(ps) < n_parents:
                ps.add(rng.below(i))
        parents.append(ps)
    return _assemble(cards, parents, seed, f"card{card}")


def random_polytree(n: int, card_hi: int = 4, max_parents: int = 3, seed: int = 3) -> FlatNetwork:
    """Random polytree (singly connected DAG, nodes may have several parents): BP is exact
    there.  Edges a->b (a<b) are accepted only if they join two different undirected components."""
    rng = SplitMix64(seed ^ 0x7EE)
    card = [2 + rng.below(card_hi - 1) for _ in range(n)]
    parents = [set() for _ in range(n)]
    comp = list(range(n))

    def find(v):
        while comp[v] != v:
            comp[v] = comp[comp[v]]
            v = comp[v]
        return v

    edges, tries = 0, 0
    while edges < n - 1 and tries < 200 * n:
        tries += 1
        a, b = rng.below(n), rng.below(n)
        if a == b:
            continue
        if a > b:
            a, b = b, a
        ra, rb = find(a), find(b)
        if ra == rb or len(parents[b]) >= max_parents:
            continue
        comp[ra] = rb
        parents[b].add(a)
        edges += 1
    return _assemble(card, parents, seed, f"polytree{n}")


# ---- evidence ------------------------------------------------------------------------------------
def make_evidence(net: FlatNetwork, n_cases: int, *, exact_k: int | None = None, p: float = 0.10,
                  seed: int = EVIDENCE_SEED, case_offset: int = 0, soft: bool = False) -> EvidenceBatch:
    """Hard evidence per case: either exactly ``exact_k`` distinct nodes (cfg 2: 4) or every node
    independently with probability ``p``; the observed state is uniform.  ``soft=True`` turns each
    entry into a strictly positive random row instead of a one-hot (tests only)."""
    n = net.n_nodes
    case = np.arange(case_offset, case_offset + n_cases, dtype=np.uint64)
    if exact_k is not None:
        k = min(exact_k, n)
        chosen = np.empty((n_cases, k), dtype=np.int64)
        for d in range(k):
            idx = (counter_u64(seed, case, draw=d) % np.uint64(n - d)).astype(np.int64)
            if d:
                prev = np.sort(chosen[:, :d], axis=1)
                for t in range(d):
                    idx += (idx >= prev[:, t]).astype(np.int64)
            chosen[:, d] = idx
        chosen.sort(axis=1)
        ev_node = chosen.reshape(-1).astype(np.int32)
        ev_off = np.arange(n_cases + 1, dtype=np.int64) * k
        draw_base = k
        ent_case = np.repeat(case, k)
        ent_slot = np.tile(np.arange(k, dtype=np.uint64), n_cases)
    else:
        thresh = np.uint64(int(p * 18446744073709551616.0) & _M64)
        node_ids = np.arange(n, dtype=np.uint64)
        chunks_nodes, counts = [], np.zeros(n_cases, dtype=np.int64)
        step = max(1, (1 << 24) // max(n, 1))
        for lo in range(0, n_cases, step):
            hi = min(n_cases, lo + step)
            c = case[lo:hi, None]
            with np.errstate(over="ignore"):
                key = c * np.uint64(0x100000001B3) + node_ids[None, :]
            pick = counter_u64(seed ^ 0x5EED, key.ravel()).reshape(hi - lo, n) < thresh
            counts[lo:hi] = pick.sum(axis=1)
            chunks_nodes.append(np.nonzero(pick)[1].astype(np.int32))
        ev_node = np.concatenate(chunks_nodes) if chunks_nodes else np.zeros(0, np.int32)
        ev_off = np.zeros(n_cases + 1, dtype=np.int64)
        np.cumsum(counts, out=ev_off[1:])
        ent_case = np.repeat(case, counts)
        ent_slot = ev_node.astype(np.uint64)
        draw_base = 0
    with np.errstate(over="ignore"):
        ent_key = ent_case * np.uint64(0x9E3779B1) + ent_slot * np.uint64(0x85EBCA77)
    r = net.card[ev_node].astype(np.uint64)
    if not soft:
        ev_state = (counter_u64(seed ^ 0xABCD, ent_key, draw=draw_base) % np.maximum(r, 1)).astype(np.int32)
        return EvidenceBatch(n_cases, ev_off, ev_node, ev_state)
    voff = np.zeros(ev_node.shape[0] + 1, dtype=np.int64)
    np.cumsum(r.astype(np.int64), out=voff[1:])
    vals = np.empty(int(voff[-1]), dtype=np.float64)
    rr = r.astype(np.int64)
    within = (np.arange(int(voff[-1]), dtype=np.int64) - np.repeat(voff[:-1], rr)).astype(np.uint64)
    with np.errstate(over="ignore"):
        vkey = np.repeat(ent_key, rr) * np.uint64(131) + within
    u = (counter_u64(seed ^ 0x50F7, vkey) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    vals[:] = 0.05 + 0.95 * u
    return EvidenceBatch(n_cases, ev_off, ev_node, None, voff, vals)


# ---- the same evidence, generated with torch (GPU) ---------------------------------------------------
def _s64(v: int) -> int:
    """A 64-bit pattern as the signed value torch.int64 holds."""
    v &= _M64
    return v - (1 << 64) if v >= (1 << 63) else v


def _lsr(z, k: int):
    """Logical right shift of int64 tensors (torch's >> is arithmetic)."""
    return (z >> k) & ((1 << (64 - k)) - 1)


def _mix_t(z):
    z = z ^ _lsr(z, 30)
    z = z * _s64(0xBF58476D1CE4E5B9)
    z = z ^ _lsr(z, 27)
    z = z * _s64(0x94D049BB133111EB)
    return z ^ _lsr(z, 31)


def _counter_t(seed: int, index, draw: int = 0):
    s0 = _mix_t(index ^ _s64(seed))
    return _mix_t(s0 + _s64((draw + 1) * _GAMMA))


def make_evidence_torch(net: FlatNetwork, n_cases: int, *, p: float = 0.10, seed: int = EVIDENCE_SEED,
                        case_offset: int = 0, device="cuda", chunk_elems: int = 1 << 25):
    """``make_evidence(net, n_cases, p=p, ...)`` (hard evidence, every node independently with probability
    ``p``) computed with torch integer arithmetic on ``device``: bit-identical CSR arrays, returned as
    tensors (ev_off int64, ev_node int32, ev_state int32) that are already resident where the device path
    wants them.  numpy needs ~6 s per 8 192 cases of the 10 000-node grid; this is what lets bench.py build
    the 65 536-case batch of BASELINE config 3 in seconds."""
    import torch
    n = net.n_nodes
    dev = torch.device(device)
    sign = -(1 << 63)
    thresh = _s64(int(p * 18446744073709551616.0) & _M64) ^ sign          # unsigned compare = signed compare of x ^ 2^63
    node_ids = torch.arange(n, dtype=torch.int64, device=dev)
    card = torch.from_numpy(net.card.astype(np.int64)).to(dev)
    two64_mod = torch.tensor([(1 << 64) % max(int(r), 1) for r in net.card], dtype=torch.int64, device=dev)
    step = max(1, chunk_elems // max(n, 1))
    nodes_out, states_out, counts_out = [], [], []
    for lo in range(0, n_cases, step):
        hi = min(n_cases, lo + step)
        case = torch.arange(case_offset + lo, case_offset + hi, dtype=torch.int64, device=dev)
        key = case[:, None] * 0x100000001B3 + node_ids[None, :]
        pick = (_counter_t(seed ^ 0x5EED, key) ^ sign) < thresh
        counts_out.append(pick.sum(dim=1))
        rows, cols = torch.nonzero(pick, as_tuple=True)
        ent_key = case[rows] * 0x9E3779B1 + cols * 0x85EBCA77
        u = _counter_t(seed ^ 0xABCD, ent_key, draw=0)
        r = card[cols]
        st = torch.remainder(u, r)
        st = torch.where(u < 0, torch.remainder(st + two64_mod[cols], r), st)   # u is an unsigned 64-bit value
        nodes_out.append(cols.to(torch.int32))
        states_out.append(st.to(torch.int32))
    counts = torch.cat(counts_out) if counts_out else torch.zeros(0, dtype=torch.int64, device=dev)
    ev_off = torch.zeros(n_cases + 1, dtype=torch.int64, device=dev)
    ev_off[1:] = torch.cumsum(counts, 0)
    ev_node = torch.cat(nodes_out) if nodes_out else torch.zeros(0, dtype=torch.int32, device=dev)
    ev_state = torch.cat(states_out) if states_out else torch.zeros(0, dtype=torch.int32, device=dev)
    return ev_off, ev_node, ev_state


WORKLOADS = {
    # name: (network factory, default cases, evidence kwargs, default fixed sweeps)
    "alarm37": (alarm37, 1 << 20, dict(exact_k=4), 20),
    "grid100": (lambda: grid(100), 1 << 16, dict(p=0.10), 50),
    "dag2000": (random_dag, 1 << 15, dict(p=0.10), 20),
    "card32": (high_card, 1 << 14, dict(p=0.10), 20),
}
