"""ctypes bindings of the parity checkers -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  ``port``  = oracle/_build/libbp_oracle.so (C restatement, bp_oracle.c);
``reference`` = oracle/_ref/libbnref.so (the reference's own headers compiled in place).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "_build", "libbp_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libbnref.so")
REF_PURE_SO = os.path.join(HERE, "_ref", "libbnref_pure.so")


def build(quiet: bool = True) -> None:
    """Compile the checkers (the reference flavour only where /root/reference exists)."""
    subprocess.run(["make", "-C", HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _ptr(a, ctype):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(ctype))


def _net_args(net):
    return (C.c_int32(net.n_nodes), _ptr(net.card, C.c_int32), _ptr(net.parent_off, C.c_int32),
            _ptr(net.parents, C.c_int32), _ptr(net.cpt_off, C.c_int64), _ptr(net.cpt, C.c_double))


def _ev_args(ev):
    return (C.c_int64(ev.n_cases), _ptr(ev.ev_off, C.c_int64), _ptr(ev.ev_node, C.c_int32),
            _ptr(ev.ev_state, C.c_int32) if not ev.is_soft else None,
            _ptr(ev.ev_val_off, C.c_int64) if ev.is_soft else None,
            _ptr(ev.ev_values, C.c_double) if ev.is_soft else None)


_port = None


def have_port() -> bool:
    return os.path.exists(PORT_SO)


def have_reference() -> bool:
    return os.path.exists(REF_SO)


def port_max_threads() -> int:
    global _port
    if _port is None:
        _port = C.CDLL(PORT_SO)
    return int(_port.bp_oracle_max_threads())


def run_port(net, ev, eps=1e-3, max_sweeps=0, damping=0.0, check_interval=1, threads=1, semiring=0):
    """C restatement.  Returns (marginals [B, sum r], sweeps [B], converged [B]).  semiring=1: max-product."""
    global _port
    if _port is None:
        _port = C.CDLL(PORT_SO)
    out = np.empty((ev.n_cases, net.belief_values), dtype=np.float64)
    sweeps = np.empty(ev.n_cases, dtype=np.int32)
    conv = np.empty(ev.n_cases, dtype=np.uint8)
    _port.bp_oracle_run_semiring.restype = C.c_int
    rc = _port.bp_oracle_run_semiring(*_net_args(net), *_ev_args(ev), C.c_double(eps), C.c_int32(max_sweeps),
                                      C.c_double(damping), C.c_int32(check_interval), C.c_int32(threads), C.c_int32(semiring),
                                      _ptr(out, C.c_double), _ptr(sweeps, C.c_int32), _ptr(conv, C.c_uint8))
    if rc != 0:
        raise RuntimeError("bp_oracle_run failed")
    return out, sweeps, conv


def run_port_lw(net, ev, n_samples, seed=1, case_base=0):
    """C restatement of the reference's likelihood weighting with the counter-based variates of
    csrc/bnbp_lw.cuh.  Returns (marginals [B, sum r], total weight [B])."""
    global _port
    if _port is None:
        _port = C.CDLL(PORT_SO)
    out = np.empty((ev.n_cases, net.belief_values), dtype=np.float64)
    wsum = np.empty(ev.n_cases, dtype=np.float64)
    _port.bp_oracle_lw.restype = C.c_int
    rc = _port.bp_oracle_lw(*_net_args(net), C.c_int64(ev.n_cases), _ptr(ev.ev_off, C.c_int64),
                            _ptr(ev.ev_node, C.c_int32), _ptr(ev.ev_state, C.c_int32), C.c_int64(n_samples),
                            C.c_uint64(seed), C.c_int64(case_base), _ptr(out, C.c_double), _ptr(wsum, C.c_double))
    if rc != 0:
        raise RuntimeError("bp_oracle_lw failed")
    return out, wsum


def port_make_cpt(net, samples, multiplicity=None):
    """C restatement of sampler::make_cpt (sampler.hpp:81-163) over flat arrays."""
    global _port
    if _port is None:
        _port = C.CDLL(PORT_SO)
    samples = np.ascontiguousarray(samples, dtype=np.int32)
    mult = None if multiplicity is None else np.ascontiguousarray(multiplicity, dtype=np.int64)
    out = np.empty(int(net.cpt_off[-1]), dtype=np.float64)
    _port.bp_oracle_make_cpt.restype = C.c_int
    rc = _port.bp_oracle_make_cpt(C.c_int32(net.n_nodes), _ptr(net.card, C.c_int32), _ptr(net.parent_off, C.c_int32),
                                  _ptr(net.parents, C.c_int32), _ptr(net.cpt_off, C.c_int64), _ptr(samples, C.c_int32),
                                  _ptr(mult, C.c_int64), C.c_int64(samples.shape[0]), _ptr(out, C.c_double))
    if rc != 0:
        raise RuntimeError("bp_oracle_make_cpt failed")
    return out


REF_LW_SO = os.path.join(HERE, "_ref", "libbnref_lw.so")


def have_reference_lw() -> bool:
    return os.path.exists(REF_LW_SO)


def run_reference_lw(net, ev, n_samples):
    """The reference's own likelihood_weighting (random_device-seeded: differs run to run)."""
    if REF_LW_SO not in _ref:
        _ref[REF_LW_SO] = C.CDLL(REF_LW_SO)
    lib = _ref[REF_LW_SO]
    out = np.empty((ev.n_cases, net.belief_values), dtype=np.float64)
    lib.bnref_lw.restype = C.c_int
    rc = lib.bnref_lw(*_net_args(net), C.c_int64(ev.n_cases), _ptr(ev.ev_off, C.c_int64), _ptr(ev.ev_node, C.c_int32),
                      _ptr(ev.ev_state, C.c_int32), C.c_int64(n_samples), _ptr(out, C.c_double))
    if rc != 0:
        raise RuntimeError("bnref_lw failed")
    return out


_ref = {}


def run_reference(net, ev, eps=1e-3, max_sweeps=0, pure=False):
    """The reference's own belief_propagation (oracle/_ref).  Returns
    (marginals, sweeps, converged, seconds inside operator())."""
    path = REF_PURE_SO if pure else REF_SO
    if path not in _ref:
        _ref[path] = C.CDLL(path)
    lib = _ref[path]
    out = np.empty((ev.n_cases, net.belief_values), dtype=np.float64)
    sweeps = np.empty(ev.n_cases, dtype=np.int32)
    conv = np.empty(ev.n_cases, dtype=np.uint8)
    secs = C.c_double(0.0)
    lib.bnref_run.restype = C.c_int
    rc = lib.bnref_run(*_net_args(net), *_ev_args(ev), C.c_double(eps), C.c_int32(max_sweeps),
                       _ptr(out, C.c_double), _ptr(sweeps, C.c_int32), _ptr(conv, C.c_uint8),
                       C.byref(secs))
    if rc != 0:
        raise RuntimeError("bnref_run failed")
    return out, sweeps, conv, secs.value


REF_DSC_SO = os.path.join(HERE, "_ref", "libbnref_dsc.so")


def have_reference_dsc() -> bool:
    return os.path.exists(REF_DSC_SO)


def reference_dsc_flatten(text: str):
    """The reference's own DSC loader (serializer/dsc.hpp) on `text`, read back through its graph
    accessors.  Returns (card, parent_off, parents, cpt_off, cpt) in the layout of include/bnbp.h."""
    if REF_DSC_SO not in _ref:
        _ref[REF_DSC_SO] = C.CDLL(REF_DSC_SO)
    lib = _ref[REF_DSC_SO]
    lib.bnref_dsc_flatten.restype = C.c_int
    raw = text.encode("utf-8")
    n, e, v = C.c_int32(0), C.c_int32(0), C.c_int64(0)
    lib.bnref_dsc_flatten(raw, C.byref(n), C.byref(e), C.byref(v), None, None, None, None, None)
    card = np.zeros(n.value, dtype=np.int32)
    poff = np.zeros(n.value + 1, dtype=np.int32)
    par = np.zeros(max(e.value, 1), dtype=np.int32)
    coff = np.zeros(n.value + 1, dtype=np.int64)
    cpt = np.zeros(max(v.value, 1), dtype=np.float64)
    lib.bnref_dsc_flatten(raw, C.byref(n), C.byref(e), C.byref(v), _ptr(card, C.c_int32), _ptr(poff, C.c_int32),
                          _ptr(par, C.c_int32), _ptr(coff, C.c_int64), _ptr(cpt, C.c_double))
    return card, poff, par[:e.value], coff, cpt[:v.value]


REF_SAMPLER_SO = os.path.join(HERE, "_ref", "libbnref_sampler.so")


def have_reference_sampler() -> bool:
    return os.path.exists(REF_SAMPLER_SO)


def _sampler_lib():
    if REF_SAMPLER_SO not in _ref:
        _ref[REF_SAMPLER_SO] = C.CDLL(REF_SAMPLER_SO)
    return _ref[REF_SAMPLER_SO]


def reference_make_cpt(net, samples, multiplicity=None):
    """The reference's own sampler::load_sample(table) + sampler::make_cpt (sampler.hpp:29-163, compiled in place
    over the Boost stand-ins of tests/cpp/boost).  Returns the CPT arena in the layout of include/bnbp.h."""
    lib = _sampler_lib()
    samples = np.ascontiguousarray(samples, dtype=np.int32)
    mult = None if multiplicity is None else np.ascontiguousarray(multiplicity, dtype=np.int64)
    out = np.empty(int(net.cpt_off[-1]), dtype=np.float64)
    lib.bnref_make_cpt.restype = C.c_int
    rc = lib.bnref_make_cpt(C.c_int32(net.n_nodes), _ptr(net.card, C.c_int32), _ptr(net.parent_off, C.c_int32),
                            _ptr(net.parents, C.c_int32), _ptr(net.cpt_off, C.c_int64), _ptr(samples, C.c_int32),
                            _ptr(mult, C.c_int64), C.c_int64(samples.shape[0]), _ptr(out, C.c_double))
    if rc != 0:
        raise RuntimeError(f"bnref_make_cpt failed ({rc})")
    return out


def reference_make_cpt_from_file(net, path):
    """The reference's own sample-file reader (sampler.hpp:42-76: 'count v1 .. vN' per line) + make_cpt.
    Returns (cpt, sampling_size)."""
    lib = _sampler_lib()
    out = np.empty(int(net.cpt_off[-1]), dtype=np.float64)
    size = C.c_int64(0)
    lib.bnref_make_cpt_from_file.restype = C.c_int
    rc = lib.bnref_make_cpt_from_file(C.c_int32(net.n_nodes), _ptr(net.card, C.c_int32), _ptr(net.parent_off, C.c_int32),
                                      _ptr(net.parents, C.c_int32), _ptr(net.cpt_off, C.c_int64), path.encode(),
                                      C.byref(size), _ptr(out, C.c_double))
    if rc != 0:
        raise RuntimeError(f"bnref_make_cpt_from_file failed ({rc})")
    return out, int(size.value)
