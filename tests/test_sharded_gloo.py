"""CPU, world_size 2 over gloo: the row-sharding plan, the single candidate all-gather and the merge logic of
ShardedFlatIndex.  The two device steps (local search, k-way merge) are replaced by the CPU oracle here -- the
product path itself has no CPU implementation."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from b200knn.sharded import candidate_views, pack_candidates, shard_rows, unpack_candidates


def test_shard_rows_cover_everything():
    for n in (0, 1, 7, 100, 50_000_000):
        for w in (1, 2, 4, 8):
            parts = shard_rows(n, w)
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            for (s0, c0), (s1, _) in zip(parts, parts[1:]):
                assert s0 + c0 == s1


def test_pack_unpack_roundtrip():
    vals = torch.randn(5, 7)
    idx = torch.randint(0, 1 << 40, (5, 7))
    buf = torch.cat([pack_candidates(vals, idx), pack_candidates(vals + 1, idx + 1)])
    v, i = unpack_candidates(buf, 2, 5, 7)
    assert torch.equal(v[0], vals) and torch.equal(i[1], idx + 1)


def test_candidate_views_alias_the_packed_buffer():
    vals = torch.randn(5, 7)
    idx = torch.randint(0, 1 << 40, (5, 7))
    buf = torch.zeros(3 * 35 + 1, dtype=torch.int32)
    v, i = candidate_views(buf, 5, 7)
    v.copy_(vals)
    i.copy_(idx)
    assert torch.equal(buf[:105], pack_candidates(vals, idx))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import oracle
    from b200knn.search import FlatIndex
    from b200knn.sharded import ShardedFlatIndex, candidate_views

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rs = np.random.RandomState(0)
    g = oracle.normalize(rs.standard_normal((301, 32)).astype(np.float32))
    g[17] = g[250]  # a tie across the shard boundary: must resolve to the lower global row
    q = g[:41].copy()
    start, count = shard_rows(g.shape[0], world)[rank]
    raw = (q * np.linspace(0.5, 3.0, q.shape[0], dtype=np.float32)[:, None]).astype(np.float32)   # un-normalised copies

    class OracleShard(ShardedFlatIndex):
        def __init__(self):
            self.local = FlatIndex.__new__(FlatIndex)
            self.local.metric, self.local.precision, self.local.device = "cosine", "fp32", torch.device("cpu")
            self.local.rows = torch.from_numpy(g[start:start + count])
            # the query-side normalise + cast of the CUDA index, replaced by the oracle's (row-wise, like the kernel)
            self.local.prepare_queries = lambda x: (torch.from_numpy(oracle.normalize(x.numpy())), None)
            self.group, self.world_size, self.rank = None, world, rank
            self.exchange, self._peer, self._prof = "allgather", None, None

        def _search_local(self, queries, k, self_mode, query_offset, out=None, prepared=None):
            queries = prepared[0] if prepared is not None else queries
            v, i = oracle.search(queries.numpy(), g[start:start + count], k, "cosine", self_mode, query_offset, start)
            out[0].copy_(torch.from_numpy(v))      # straight into the exchange buffer, as the CUDA search does
            out[1].copy_(torch.from_numpy(i))
            return out

        def _merge_parts(self, bufs, nq, k):
            views = [candidate_views(b, nq, k) for b in bufs]
            v, i = oracle.merge_topk(np.stack([v.numpy() for v, _ in views]), np.stack([i.numpy() for _, i in views]),
                                     "cosine")
            return torch.from_numpy(v), torch.from_numpy(i)

    shard = OracleShard()
    v, i = shard.search(torch.from_numpy(q), 10, exclude_self=True)
    # host-resident batch: every rank prepares only its slice (21 + 20 rows), the slices are all-gathered
    vh, ih = shard.search_host(torch.from_numpy(raw), 10, exclude_self=True)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), v=v.numpy(), i=i.numpy(), vh=vh.numpy(), ih=ih.numpy())
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_search_equals_single_shard(tmp_path):
    import oracle

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    rs = np.random.RandomState(0)
    g = oracle.normalize(rs.standard_normal((301, 32)).astype(np.float32))
    g[17] = g[250]
    q = g[:41].copy()
    v1, i1 = oracle.search(q, g, 10, "cosine", "exclude", 0)
    raw = (q * np.linspace(0.5, 3.0, q.shape[0], dtype=np.float32)[:, None]).astype(np.float32)
    v2, i2 = oracle.search(oracle.normalize(raw), g, 10, "cosine", "exclude", 0)
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        assert np.array_equal(z["i"], i1) and np.array_equal(z["v"], v1)
        assert np.array_equal(z["ih"], i2) and np.array_equal(z["vh"], v2)      # search_host == one-shard search


def _gather_worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from b200knn.fullrank import gather_query_sharded, query_slice

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nq = 11                                            # uneven slices: 5 + 6
    full = {"ap": torch.arange(nq, dtype=torch.float64) * 0.5, "hits": torch.arange(nq * 3, dtype=torch.int32).view(nq, 3)}
    s, e = query_slice(nq, world, rank)
    out = gather_query_sharded({k: v[s:e].clone() for k, v in full.items()}, nq)
    ok = all(torch.equal(out[k], full[k]) for k in full)
    with open(os.path.join(out_dir, f"g{rank}.txt"), "w") as fh:
        fh.write("ok" if ok else "bad")
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_query_sharded_statistics_gather_in_query_order(tmp_path):
    """The multi-GPU form of the full-ranking metrics shards the QUERIES; the per-query statistics of uneven slices must
    come back in the original order on every rank (gloo, CPU tensors)."""
    world = 2
    mp.spawn(_gather_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"g{r}.txt").read() == "ok"
