import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import unicorn_oracle as orc
from unicorn_b200.weights import make_state_dict
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    if os.path.exists(p): print(p, open(p).read().strip())
name = "unicorn_track_large"
sd = make_state_dict(name, 0)
img = torch.rand(1, 3, 800, 1280) * 255
for nt in (8, 16, 32, 64):
    torch.set_num_threads(nt)
    with torch.no_grad():
        t0 = time.perf_counter()
        orc.convnext_features(img, sd, orc.CONFIGS[name])
        print("threads", nt, "convnext_features", round(time.perf_counter() - t0, 2), "s", flush=True)
