"""Workloads for the round-2 ncu captures of the secondary kernels (development helper):
   merge   : 8192 x 6.25 M x 512 bf16 top-100 (the per-GPU shard of C5 at N=8) -> merge_units_kernel
   hamming : 128 queries x 10 M 64-bit codes, top-100, popcount and +-1 tensor-core forms
   fullrank: evaluate_map_embeddings at 30 k x 1024 (rank_positives_kernel, AP kernels, dense FFMA)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200knn
what = sys.argv[1]
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(1)
if what == "merge":
    g = b200knn.normalize(torch.randn((6_250_000, 512), generator=gen, device=dev), out_dtype=torch.bfloat16)
    q = b200knn.normalize(torch.randn((8192, 512), generator=gen, device=dev), out_dtype=torch.bfloat16)
    ix = b200knn.FlatIndex(512, "cosine", "bf16").adopt(g)
    for _ in range(2):
        ix.search(q, 100)
elif what == "hamming":
    g = torch.randint(-2**62, 2**62, (10_000_000, 1), generator=gen, device=dev, dtype=torch.int64)
    q = torch.randint(-2**62, 2**62, (128, 1), generator=gen, device=dev, dtype=torch.int64)
    for m in ("popc", "mma"):
        for _ in range(2):
            b200knn.search_hamming(q, g, 100, packed=True, bits=64, method=m)
elif what == "fullrank":
    from oracle import synth
    n, d = 30_000, 1024
    ml = torch.from_numpy(synth.multihot(n, seed=3)).cuda()
    emb = b200knn.normalize(torch.randn((n, d), generator=gen, device=dev))
    for _ in range(2):
        print(b200knn.metrics.evaluate_map_embeddings(emb, ml, 0.4))
torch.cuda.synchronize()
