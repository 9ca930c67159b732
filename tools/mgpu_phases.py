"""torchrun: per-phase wall times of the config-3 e2e call on N ranks (developer tool)."""
import os, sys, time, warnings
import numpy as np
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-gstat_b200"))
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
import skgstat_b200 as skg
from skgstat_b200 import engine as E
from skgstat_b200.data import gaussian_field, random_points
warnings.simplefilter("ignore")
c = random_points(200000, 10000.0, 2, seed=0); v = gaussian_field(c, 1000.0, seed=0)
mk = lambda: skg.Variogram(c.copy(), v.copy(), estimator="cressie", n_lags=25, maxlag="median", fit_method=None).experimental
for _ in range(3): mk()
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(5): mk()
torch.cuda.synchronize()
if rank == 0: print("e2e %.2f ms on %d ranks" % ((time.perf_counter() - t0) / 5 * 1e3, dist.get_world_size()))
def wrap(cls, name):
    f = getattr(cls, name)
    def g(self, *a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = f(self, *a, **k)
        torch.cuda.synchronize()
        if rank == 0: print("   [call] %-22s %.3f ms (n=%d)" % (name, (time.perf_counter() - t0) * 1e3, getattr(self, 'n', 0)))
        return r
    setattr(cls, name, g)
for nm in ("__init__", "set_values", "hist", "sq_order_stats", "hist_counts", "sq_cell_counts"):
    wrap(E.PairEngine, nm)
t0 = time.perf_counter(); mk(); torch.cuda.synchronize()
if rank == 0: print("instrumented total %.2f ms" % ((time.perf_counter() - t0) * 1e3))
dist.destroy_process_group()
