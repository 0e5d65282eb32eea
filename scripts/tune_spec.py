"""Tuning sweep of the network-specialised kernel's knobs (cases per thread, min blocks per SM,
load look-ahead).  `precompile` runs here without a GPU (fills the in-tree cubin cache in
parallel); `run` executes bench.py per combination on the GPU box and prints one line each."""
import json
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMBOS = {
    "fp64": [(1, 3, 1), (1, 4, 1), (1, 2, 1), (1, 3, 0), (1, 3, 2), (2, 2, 1), (2, 1, 1), (2, 3, 1)],
    "fp32": [(2, 4, 1), (4, 2, 1), (4, 3, 1), (2, 3, 1), (1, 4, 1), (2, 4, 2)],
}


def env_for(vec, minb, ahead):
    e = dict(os.environ)
    e.update(BNBP_SPEC_VEC=str(vec), BNBP_SPEC_MINB=str(minb), BNBP_SPEC_AHEAD=str(ahead))
    return e


def precompile():
    def one(job):
        prec, (vec, minb, ahead) = job
        code = (f"import sys; sys.path.insert(0, {ROOT!r}); from bayesiannetwork_b200 import engine, synth; "
                f"engine.precompile(synth.alarm37(), {prec!r}, 0b11001)")
        r = subprocess.run([sys.executable, "-c", code], env=env_for(vec, minb, ahead), capture_output=True, text=True)
        return job, r.returncode, r.stderr[-300:]
    jobs = [(p, c) for p, cs in COMBOS.items() for c in cs]
    with ThreadPoolExecutor(8) as ex:
        for job, rc, err in ex.map(one, jobs):
            print(job, "ok" if rc == 0 else err, flush=True)


def run():
    for prec, cs in COMBOS.items():
        for (vec, minb, ahead) in cs:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--precision", prec, "--no-cpu", "--no-e2e",
                                "--steps", "3", "--warmup", "3", "--specialize", "always"],
                               env=env_for(vec, minb, ahead), capture_output=True, text=True)
            try:
                j = json.loads(r.stdout.strip().splitlines()[-1])
                print(f"{prec} vec={vec} minb={minb} ahead={ahead}: {j['roofline']['ms_per_launch']:.3f} ms/sweep "
                      f"frac={j['roofline']['frac']:.3f} value={j['value']:.4g}", flush=True)
            except Exception:
                print(prec, vec, minb, ahead, "FAILED", r.stderr[-400:], flush=True)


if __name__ == "__main__":
    {"precompile": precompile, "run": run}[sys.argv[1]]()
