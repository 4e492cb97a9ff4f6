"""End-to-end timing of the full-ranking metrics at the NIH scale (development helper)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200knn
from b200knn import fullrank as FR
from oracle import synth
n, d = int(os.environ.get("N", 112000)), 1024
ml = synth.multihot(n, seed=3)
mld = torch.from_numpy(ml).cuda()
gen = torch.Generator(device="cuda").manual_seed(7)
mu = torch.randn((14, d), generator=gen, device="cuda")
emb = b200knn.normalize(mld @ mu / mld.sum(1, keepdim=True).clamp(min=1) + torch.randn((n, d), generator=gen, device="cuda"))
lab = torch.randint(0, 3, (n,), device="cuda")
M = b200knn.metrics
def t(name, fn, reps=2):
    for r in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); v = fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{name}: {dt:.3f} s -> {v if not isinstance(v, dict) else {k: round(x, 4) for k, x in v.items()}}", flush=True)
for budget in [int(x) for x in os.environ.get("BUDGETS", "3,8").split(",")]:
    FR._CHUNK_BYTES = budget << 30
    print("budget GiB", budget, "rows/chunk ties", FR.chunk_rows(n, True, FR._CHUNK_BYTES, "cuda"), "plain", FR.chunk_rows(n, False, FR._CHUNK_BYTES, "cuda"))
    t("evaluate_map_embeddings (D8, sklearn AP)", lambda: M.evaluate_map_embeddings(emb, mld, 0.4))
    t("compute_map_multilabel_from_embeddings (D4)", lambda: M.compute_map_multilabel_from_embeddings(emb, mld, 0.5))
    t("_compute_single_label_retrieval_metrics (D6)", lambda: M._compute_single_label_retrieval_metrics(emb, lab))
    t("compute_map_from_embeddings (D2)", lambda: M.compute_map_from_embeddings(emb, lab, [1, 5, 10], metric="cosine")[0])
# dense only
q = emb[:1184]
from b200knn.search import _scores_dense_prepared
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sc = _scores_dense_prepared(q, None, emb, None, "cosine", "exclude", 0); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
print(f"dense 1184 x {n}: {dt*1e3:.2f} ms -> {2*1184*n*d/dt/1e12:.1f} TFLOP/s; all rows {dt*n/1184:.3f} s")
for prec in ("bf16x3", "bf16"):
    t(f"evaluate_map_embeddings precision={prec}", lambda: M.evaluate_map_embeddings(emb, mld, 0.4, precision=prec))
    t(f"compute_map_multilabel precision={prec}", lambda: M.compute_map_multilabel_from_embeddings(emb, mld, 0.5, precision=prec))
