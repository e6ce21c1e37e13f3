import sys, os, time
sys.path.insert(0, os.getcwd())
import torch
from torch.profiler import profile, ProfilerActivity
from open3dsot_b200.config import load_config
from open3dsot_b200.datasets.device_sampler import DeviceSiameseSampler
from open3dsot_b200.datasets.synthetic import synthetic_sequence
cfg = load_config("cfgs/BAT_Car.yaml", {"batch_size": 48})
tr = [synthetic_sequence(n_frames=8, n_points=60000, seed=20260924 + i) for i in range(6)]
smp = DeviceSiameseSampler(tr, cfg, "cuda", seed=1, use_graph="--eager" not in sys.argv)
for _ in range(3): smp.next_batch()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): smp.next_batch()
t_host = (time.perf_counter() - t0) / 10
torch.cuda.synchronize()
t_all = (time.perf_counter() - t0) / 10
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): smp.next_batch()
    torch.cuda.synchronize()
agg = {}
for e in prof.events():
    if e.device_type.name != "CUDA": continue
    a = agg.setdefault(e.name[:70], [0, 0.0]); a[0] += 1; a[1] += e.device_time
tot = sum(v[1] for v in agg.values())
print(f"host enqueue {t_host*1e3:.2f} ms/batch, wall {t_all*1e3:.2f} ms/batch, kernel time {tot/3e3:.2f} ms/batch, launches {sum(v[0] for v in agg.values())/3:.0f}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"{v[1]/3e3:7.3f} ms {v[0]/3:6.1f}  {k}")
