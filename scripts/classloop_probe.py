"""A/B of the kernel families on cfg 3 (the 100 x 100 grid) -- GPU box only.  One process, one evidence batch resident on
the device, one handle per variant: the generic sweep kernel, then the class-looped specialised kernel at several
(cases per thread, resident blocks per SM) settings.  Prints one JSON line per variant: ms per step, the sweep kernel's
time per sweep and its fraction of the measured HBM copy rate, plus the checks -- class-looped fp64 variants equal each
other bit for bit, equal the generic kernel to rounding (whole batch, compared on the device) and the oracle port on a
sample.  The kernels of every setting are precompiled by ``--precompile`` (CPU box, no GPU).

    python scripts/classloop_probe.py --precompile            # here
    gpurun -- 'python scripts/classloop_probe.py --cases 65536 > gpurun_out/probe.jsonl'
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# (name, precision, specialize, env)
VARIANTS = [
    ("generic", "fp64", "never", {}),
    ("class_v1_b4", "fp64", "auto", {"BNBP_SPEC_VEC": "1", "BNBP_SPEC_MINB": "4"}),
    ("class_v1_b3", "fp64", "auto", {"BNBP_SPEC_VEC": "1", "BNBP_SPEC_MINB": "3"}),
    ("class_v2_b2", "fp64", "auto", {"BNBP_SPEC_VEC": "2", "BNBP_SPEC_MINB": "2"}),
    ("class_v1_b5", "fp64", "auto", {"BNBP_SPEC_VEC": "1", "BNBP_SPEC_MINB": "5"}),
    ("generic_fp32", "fp32", "never", {}),
    ("class32_v1_b5", "fp32", "auto", {"BNBP_SPEC_VEC": "1", "BNBP_SPEC_MINB": "5"}),
    ("class32_v2_b4", "fp32", "auto", {"BNBP_SPEC_VEC": "2", "BNBP_SPEC_MINB": "4"}),
    ("class32_v1_b8", "fp32", "auto", {"BNBP_SPEC_VEC": "1", "BNBP_SPEC_MINB": "8"}),
    ("class32_v4_b3", "fp32", "auto", {"BNBP_SPEC_VEC": "4", "BNBP_SPEC_MINB": "3"}),
    ("class32_v4_b4", "fp32", "auto", {"BNBP_SPEC_VEC": "4", "BNBP_SPEC_MINB": "4"}),
    ("class32_v4_b2", "fp32", "auto", {"BNBP_SPEC_VEC": "4", "BNBP_SPEC_MINB": "2"}),
    ("class32_v2_b5", "fp32", "auto", {"BNBP_SPEC_VEC": "2", "BNBP_SPEC_MINB": "5"}),
    ("class32_v2_b6", "fp32", "auto", {"BNBP_SPEC_VEC": "2", "BNBP_SPEC_MINB": "6"}),
    ("class32_v2_b3", "fp32", "auto", {"BNBP_SPEC_VEC": "2", "BNBP_SPEC_MINB": "3"}),
]
KNOBS = ("BNBP_SPEC_VEC", "BNBP_SPEC_MINB")


def set_env(env):
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update(env)


def precompile():
    from concurrent.futures import ProcessPoolExecutor
    with ProcessPoolExecutor(8) as ex:
        for r in ex.map(_precompile_one, [v for v in VARIANTS if v[2] != "never"]):
            print(r, flush=True)


def _precompile_one(v):
    from bayesiannetwork_b200 import engine, synth
    set_env(v[3])
    engine.precompile(synth.grid(100), v[1], 0b11001)        # fixed-count runs: variants 0, 3, 4
    return v[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precompile", action="store_true")
    ap.add_argument("--cases", default=str(1 << 16), help="batch size, or a comma-separated list of them")
    ap.add_argument("--sweeps", type=int, default=50)
    ap.add_argument("--only", default="", help="comma-separated variant names")
    args = ap.parse_args()
    if args.precompile:
        precompile()
        return
    for n in [int(c) for c in str(args.cases).split(",")]:
        run(args, n)


def run(args, n):
    import numpy as np
    import torch
    from bayesiannetwork_b200 import synth
    from bayesiannetwork_b200.engine import BeliefPropagation
    from bayesiannetwork_b200.flat import EvidenceBatch
    from oracle import oracle

    dev = torch.device("cuda:0")
    net = synth.grid(100)
    sweeps = args.sweeps
    S, V = net.state_values, net.belief_values
    peak = 6544.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    d_off, d_node, d_state = synth.make_evidence_torch(net, n, device=dev, p=0.10)
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream().cuda_stream
    # oracle sample: 8 evenly spaced cases
    idx = np.unique(np.linspace(0, n - 1, 8).astype(np.int64))
    offs = d_off.cpu().numpy()
    cases = []
    for c in idx:
        a, b = int(offs[c]), int(offs[c + 1])
        cases.append(dict(zip(d_node[a:b].cpu().numpy().tolist(), d_state[a:b].cpu().numpy().tolist())))
    sample = EvidenceBatch.from_cases(net, cases)
    if not oracle.have_port():
        oracle.build()
    om, _, _ = oracle.run_port(net, sample, eps=0.0, max_sweeps=sweeps, threads=0)
    only = [s for s in args.only.split(",") if s]
    ref = {}                                                 # precision -> (generic output, first class-looped output)
    for name, prec, spec, env in VARIANTS:
        if only and name not in only:
            continue
        set_env(env)
        tdtype = torch.float64 if prec == "fp64" else torch.float32
        tsize = 8 if prec == "fp64" else 4
        out = torch.empty((n, V), dtype=tdtype, device=dev)
        line = {"variant": name, "precision": prec, "cases": n, "sweeps": sweeps, "env": env}
        if prec == "fp32":
            ref.pop("fp64", None)                            # the fp64 outputs (10 GB each) are no longer compared
        bp = None
        try:
            bp = BeliefPropagation(net, prec, device=0, specialize=spec)
            bp.run_device(n, d_off, d_node, d_state, out, epsilon=0.0, max_sweeps=sweeps, stream=stream)   # warm-up: allocates
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            bp.run_device(n, d_off, d_node, d_state, out, epsilon=0.0, max_sweeps=sweeps, stream=stream)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            st = bp.stats()
            sweep_ms = st["last_sweep_ms"] / max(1, st["last_sweep_launches"])
            resident = min(n, st["resident_cases"])
            line.update(ms_per_step=ms, case_sweeps_per_s=n * sweeps / (ms * 1e-3), step_hbm_frac=2.0 * S * tsize * n * sweeps / (ms * 1e-3) / 1e9 / peak,
                        sweep_ms=sweep_ms, sweep_hbm_frac=2.0 * S * tsize * resident / (sweep_ms * 1e-3) / 1e9 / peak,
                        resident_cases=int(st["resident_cases"]), class_count=int(st["spec_class_count"]),
                        specialised=int(st["last_specialised"]), cases_per_tile=int(st["cases_per_tile"]),
                        spec_compile_ms=st["spec_compile_ms"], kernel_launches=int(st["last_kernel_launches"]))
            bp.close()                                       # frees the state arena (104 GB in fp64) before the comparisons
            bp = None
            rows = out[torch.from_numpy(idx).to(dev)].double().cpu().numpy()
            rtol, atol = (1e-9, 1e-12) if prec == "fp64" else (1e-5, 1e-7)
            err = np.abs(rows - om)
            line["oracle_max_err_over_bound"] = float(np.nanmax(err / (rtol * np.maximum(np.abs(rows), np.abs(om)) + atol)))
            g, c = ref.get(prec, (None, None))
            if spec == "never":
                ref[prec] = (out, c)
            else:
                if g is not None:                            # whole batch, in blocks of rows (no 10 GB temporaries)
                    line["max_abs_diff_vs_generic"] = max(float((out[i:i + 4096] - g[i:i + 4096]).abs().max()) for i in range(0, n, 4096))
                if c is not None:
                    line["equals_first_class_variant_bitwise"] = bool(torch.equal(out, c))
                else:
                    ref[prec] = (g, out)
        except Exception as e:                               # a variant that fails must not take the others with it
            line["error"] = repr(e)[:500]
        finally:
            if bp is not None:
                bp.close()
        print(json.dumps(line), flush=True)
        del out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
