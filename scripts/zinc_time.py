"""Forward time of the ZINC-shaped GINE-KAGIN batch (development aid; KAGNN_LIB selects the library variant)."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch
import bench_configs as B
from kagnn_b200 import models_regr
gen = torch.Generator().manual_seed(12345)
dev = B.dev
nn_, batch, ei = B.batch_of_graphs(1024, 23.15, 50, gen)
xz = torch.randint(0, 28, (nn_, 1), generator=gen); ea = torch.randint(1, 4, (ei.size(1),), generator=gen)
mz = models_regr.KAGIN(1, 1, 4, 128, 2, 5, 3, 1, 0.0, True).eval().to(dev)
dz = B.Data(xz.to(dev), ei.to(dev), batch.to(dev), ea.to(dev), 1024)
print(os.environ.get("KAGNN_LIB", "default"), "zinc ms", [round(B.timeit(lambda: mz(dz)), 4) for _ in range(3)])
