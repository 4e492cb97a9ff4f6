"""evaluate_map_embeddings at the NIH scale with the queries sharded over the ranks (torchrun; development helper)."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200knn
from oracle import synth
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
n, d = 112_000, 1024
ml = torch.from_numpy(synth.multihot(n, seed=3)).to(dev)
gen = torch.Generator(device=dev).manual_seed(7)       # the same embeddings on every rank
mu = torch.randn((14, d), generator=gen, device=dev)
emb = b200knn.normalize(ml @ mu / ml.sum(1, keepdim=True).clamp(min=1) + torch.randn((n, d), generator=gen, device=dev))
lab = torch.randint(0, 3, (n,), generator=gen, device=dev)
M = b200knn.metrics
for name, fn in (("evaluate_map_embeddings (D8)", lambda: M.evaluate_map_embeddings(emb, ml, 0.4, distributed=True)),
                 ("compute_map_multilabel_from_embeddings (D4)", lambda: M.compute_map_multilabel_from_embeddings(emb, ml, 0.5, distributed=True)),
                 ("_compute_single_label_retrieval_metrics (D6)", lambda: M._compute_single_label_retrieval_metrics(emb, lab, distributed=True)["mAP"])):
    for rep in range(2):
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        v = fn()
        torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
    if rank == 0:
        print(f"{world} GPUs, 112000 x 1024: {name}: {dt:.3f} s -> {v}", flush=True)
dist.destroy_process_group()
