"""Multi-GPU parity (needs >= 2 CUDA devices; skipped on a one-GPU box): the row-sharded search over NCCL ranks with
both candidate exchanges -- NCCL all-gather merged in place, and the merge kernel reading the peers' symmetric-memory
buffers over NVLink -- must return, on every rank, exactly what one GPU returns over the whole gallery."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist

    import b200knn
    from b200knn.sharded import ShardedFlatIndex

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(5)                                   # same data on every rank
    failures = []
    for precision, nq, ng, d, k, metric in (("bf16", 300, 70_001, 128, 100, "cosine"), ("fp32", 64, 9_000, 96, 10, "l2"),
                                            ("bf16", 1, 50_000, 768, 100, "cosine"),
                                            ("fp32-tensor", 300, 40_000, 128, 50, "cosine")):
        os.environ.pop("KNN_EXACT_ENGINE", None)
        if precision == "fp32-tensor":      # the tensor-core exact engine on every shard (split filter + proof)
            os.environ["KNN_EXACT_ENGINE"] = "tensor"
            precision = "fp32"
        g = b200knn.normalize(torch.randn((ng, d), generator=gen, device=dev))
        q = b200knn.normalize(torch.randn((nq, d), generator=gen, device=dev))
        g[17] = g[ng - 5]                                # a tie across the shard boundary
        want_v, want_i = b200knn.search(q, g, k, metric, precision=precision)
        for exchange in ("allgather", "peer"):
            sh = ShardedFlatIndex.from_full(g, metric, precision, exchange=exchange)
            for rep in range(3):                         # the peer exchange alternates between two slots
                v, i = sh.search(q, k)
                if not (torch.equal(i, want_i) and torch.equal(v, want_v)):
                    failures.append((precision, exchange, rep, int((i != want_i).sum())))
            # a HOST batch: every rank copies / prepares its slice only, the prepared slices are all-gathered
            sh.profile(True)
            v, i = sh.search_host(q.cpu().pin_memory(), k)
            if not (torch.equal(i, want_i) and torch.equal(v, want_v)):
                failures.append((precision, exchange, "search_host", int((i != want_i).sum())))
            prof = sh.profile_read()
            if len(prof) != 1 or min(prof[0]) < 0.0:
                failures.append((precision, exchange, "profile", prof))
    # pipelined peer exchange: searches enqueued back to back, the exchange of each on a side stream; alternating query
    # batches (a stale or early-rewritten slot would return the other batch's answer) and a rank that lags by a random
    # amount before every search (exercises the four-slot protocol)
    os.environ.pop("KNN_EXACT_ENGINE", None)
    g = b200knn.normalize(torch.randn((60_001, 128), generator=gen, device=dev))
    qa = b200knn.normalize(torch.randn((300, 128), generator=gen, device=dev))
    qb = b200knn.normalize(torch.randn((300, 128), generator=gen, device=dev))
    want = [b200knn.search(x, g, 100, "cosine", precision="bf16") for x in (qa, qb)]
    sh = ShardedFlatIndex.from_full(g, "cosine", "bf16", exchange="peer", pipeline=True)
    import random
    lag = random.Random(100 + rank)
    pend = []
    for step in range(24):
        if lag.random() < 0.5:
            torch.cuda._sleep(int(lag.random() * 3e6))      # up to ~1.5 ms of skew on this rank's search stream
        pend.append((step & 1, sh.search_async(qa if step & 1 == 0 else qb, 100)))
        if step % 5 == 4:                                     # take results late and in bursts
            while pend:
                which, p = pend.pop(0)
                v, i = p.result()
                if not (torch.equal(i, want[which][1]) and torch.equal(v, want[which][0])):
                    failures.append(("pipelined", step, which, int((i != want[which][1]).sum())))
    while pend:
        which, p = pend.pop(0)
        v, i = p.result()
        if not (torch.equal(i, want[which][1]) and torch.equal(v, want[which][0])):
            failures.append(("pipelined-tail", which, int((i != want[which][1]).sum())))
    # results copied to the host on the side stream (PendingSearch.to_host), the next searches already enqueued
    host_bufs = [(torch.empty((300, 100), dtype=torch.float32).pin_memory(),
                  torch.empty((300, 100), dtype=torch.int64).pin_memory()) for _ in range(6)]
    copies = []
    for step in range(6):
        if lag.random() < 0.5:
            torch.cuda._sleep(int(lag.random() * 3e6))
        p = sh.search_async(qa if step & 1 == 0 else qb, 100)
        copies.append((step & 1, p, p.to_host(*host_bufs[step])))
    for step, (which, p, ev) in enumerate(copies):
        ev.synchronize()
        hv, hi = host_bufs[step]
        if not (torch.equal(hi, want[which][1].cpu()) and torch.equal(hv, want[which][0].cpu())):
            failures.append(("pipelined-to-host", step, which))
    v, i = sh.search_host(qb.cpu().pin_memory(), 100)         # the synchronous entry points on a pipelined index
    if not (torch.equal(i, want[1][1]) and torch.equal(v, want[1][0])):
        failures.append(("pipelined", "search_host"))
    # full-ranking metrics with the QUERIES sharded over the ranks: same value as one process, on every rank
    os.environ.pop("KNN_EXACT_ENGINE", None)
    from oracle import synth
    M = b200knn.metrics
    n = 2001                                             # uneven query slices
    ml = torch.from_numpy(synth.multihot(n, seed=8)).to(dev)
    emb = torch.from_numpy(synth.labelset_clustered(synth.multihot(n, seed=8), 64, 9, 1.0)).to(dev)
    lab = torch.from_numpy(synth.clustered(n, 8, 4, seed=10)[1]).to(dev)
    for name, fn in (("evaluate_map", lambda d: M.evaluate_map_embeddings(emb, ml, 0.4, distributed=d)),
                     ("map_multilabel", lambda d: M.compute_map_multilabel_from_embeddings(emb, ml, 0.5, distributed=d)),
                     ("single_label", lambda d: M._compute_single_label_retrieval_metrics(emb, lab, distributed=d)),
                     ("multilabel", lambda d: M._compute_multilabel_retrieval_metrics(emb, ml, distributed=d)),
                     ("compute_map", lambda d: M.compute_map_from_embeddings(emb, lab, [1, 5, 10], "cosine",
                                                                             distributed=d)[0])):
        if fn(True) != fn(False):
            failures.append(("distributed", name))
    torch.cuda.synchronize()
    with open(os.path.join(out_dir, f"r{rank}.txt"), "w") as fh:
        fh.write(repr(failures))
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sharded_search_over_two_gpus_both_exchanges(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"r{r}.txt").read() == "[]"
