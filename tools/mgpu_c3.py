"""BASELINE.json config 3 (N=200k, cressie + dowd) under torchrun: e2e time of the sharded path."""
import os, sys, time, warnings
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-gstat_b200"))
import skgstat_b200 as skg
from skgstat_b200.data import gaussian_field, random_points
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
warnings.simplefilter("ignore")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
c = random_points(n, 10000.0, 2, seed=0); v = gaussian_field(c, 1000.0, seed=0)
P = n * (n - 1) // 2
for est in ("matheron", "cressie", "dowd"):
    f = lambda: skg.Variogram(c, v, estimator=est, n_lags=25, maxlag="median", fit_method=None,
                              process_group=dist.group.WORLD).experimental
    f(); torch.cuda.synchronize(); dist.barrier()
    ts = []
    for _ in range(2):
        t0 = time.perf_counter(); e = f(); torch.cuda.synchronize(); dist.barrier(); ts.append(time.perf_counter() - t0)
    if rank == 0:
        print("C3 N=%d %-9s x%d GPUs e2e %.1f ms  %.3e pairs/s  exp[0]=%.9g" % (n, est, dist.get_world_size(), min(ts) * 1e3, P / min(ts), e[0]))
dist.destroy_process_group()
