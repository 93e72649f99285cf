import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_generative_b200 import models
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = models.ImageGPT(3, 3, 32, 24, 8, 512).to(dev).eval()
m(torch.rand(16, 3, 32, 32, device=dev))
def t(label):
    torch.cuda.synchronize(); t0 = time.perf_counter(); m.sample(n_samples=16); torch.cuda.synchronize()
    print(label, f"{(time.perf_counter()-t0)*1e3:.0f} ms", flush=True)
t("incremental (capture)     ")
t("incremental (cached graph)")
m._incremental_sampling = False
m._sample_with_graphs = False
t("eager, row-truncated      ")
m._row_truncated_sampling = False
t("eager, full forward       ")

for k, v in m.__dict__.get("_samplers", {}).items():
    print("sampler", k, "graph:", type(v["graph"]).__name__, v.get("graph_error", "")[:300])
