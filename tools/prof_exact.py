"""Workload for the ncu captures of the exact engine at BASELINE config 3: two-product split filter + re-scoring."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200knn
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(1)
g = b200knn.normalize(torch.randn((112000, 1024), generator=gen, device=dev))
q = b200knn.normalize(torch.randn((25000, 1024), generator=gen, device=dev))
ix = b200knn.FlatIndex(1024, "cosine", "fp32").adopt(g)
os.environ["KNN_EXACT_ENGINE"] = "tensor"
for _ in range(2):
    v, i = ix.search(q, 50)
torch.cuda.synchronize()
