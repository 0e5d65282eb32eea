"""Where do the host-side stalls of the host-buffer call come from?  cgroup CPU throttling counters and
per-call wall time, with the interpreter doing nothing else."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def cg(name):
    for base in ("/sys/fs/cgroup", "/sys/fs/cgroup/cpu"):
        p = os.path.join(base, name)
        if os.path.exists(p):
            return open(p).read().strip().replace("\n", " | ")
    return "n/a"

print("cpu.max:", cg("cpu.max"), "| cpu.stat:", cg("cpu.stat"))
print("nproc", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "loadavg", open("/proc/loadavg").read().strip())
# pure host sleep/wake jitter: 200 x 5 ms sleeps
ts = []
for _ in range(200):
    t0 = time.perf_counter(); time.sleep(0.005); ts.append(1e3 * (time.perf_counter() - t0))
print("sleep(5ms) wall: median %.2f max %.2f ms, >10ms: %d" % (sorted(ts)[100], max(ts), sum(t > 10 for t in ts)))
# pure host compute jitter: 200 x ~5 ms busy loops
import numpy as np
a = np.random.rand(1 << 18)
ts = []
for _ in range(300):
    t0 = time.perf_counter(); a.sort(); a[::-1].sort(); ts.append(1e3 * (time.perf_counter() - t0))
print("busy loop wall: median %.2f max %.2f ms, >3x median: %d" % (sorted(ts)[150], max(ts), sum(t > 3 * sorted(ts)[150] for t in ts)))
print("cpu.stat after:", cg("cpu.stat"))
