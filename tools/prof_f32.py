"""Workload for the ncu capture of search_f32_kernel (dense mode, the distance block of the full-ranking metrics):
4736 x 112 000 x 1024 fp32, plus a CUDA-event timing of the same call."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200knn
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(1)
g = b200knn.normalize(torch.randn((112000, 1024), generator=gen, device=dev))
q = b200knn.normalize(torch.randn((4736, 1024), generator=gen, device=dev))
os.environ["KNN_EXACT_ENGINE"] = "ffma"
for _ in range(2):
    s = b200knn.scores_dense(q, g, "cosine")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    s = b200knn.scores_dense(q, g, "cosine")
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print({"ms": ms, "tflops": 2 * 4736 * 112000 * 1024 / ms / 1e9})
