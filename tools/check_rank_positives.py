"""GPU check of knn_rank_of_positives against oracle/rank_oracle.py (development helper; the tests proper are in tests/)."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200knn
from b200knn import fullrank as FR
from oracle import rank_oracle as RO

def one(nq, ng, mode, ties, drop, quant=None, seed=0, largest=True):
    rs = np.random.RandomState(seed)
    s = rs.standard_normal((nq, ng)).astype(np.float32)
    if quant: s = np.round(s * quant) / quant
    if mode == 0:
        ql = rs.randint(0, 3, nq); gl = rs.randint(0, 3, ng)
        rel = gl[None, :] == ql[:, None]
    else:
        ql = rs.randint(1, 64, nq); gl = rs.randint(1, 64, ng)
        inter = np.array([[bin(a & b).count("1") for b in gl] for a in ql]); uni = np.array([[bin(a | b).count("1") for b in gl] for a in ql])
        rel = (inter.astype(np.float32) / (uni.astype(np.float32) + np.float32(1e-8))) > np.float32(0.4)
    sd = torch.from_numpy(s).cuda()
    rp = FR.rank_of_positives(sd, mode, torch.from_numpy(ql).cuda(), torch.from_numpy(gl).cuda(), largest_first=largest,
                              jaccard_threshold=0.4, self_offset=3, drop_self=drop, ties=ties)
    torch.cuda.synchronize()
    pr, npos, nrk = rp["pos_ranks"].cpu().numpy(), rp["npos"].cpu().numpy(), rp["nranked"].cpu().numpy()
    kap = (1, 5, 10)
    st = {k: v.cpu().numpy() for k, v in FR.ap_from_ranks(rp, kap, drop).items()}
    aps = FR.ap_sklearn_from_ranks(rp).cpu().numpy() if ties else None
    bad = 0
    for q in range(nq):
        dropped = np.zeros(ng, bool)
        r = rel[q].copy()
        if drop and 0 <= q + 3 < ng:
            dropped[q + 3] = True; r[q + 3] = False
        pos, ge, tg, n, ngr = RO.rank_of_positives(s[q], r, largest, dropped)
        ok = npos[q] == len(pos) and nrk[q] == n and np.array_equal(pr[q, :len(pos)], pos)
        if ties:
            ok = ok and np.array_equal(rp["pos_ge"][q, :len(pos)].cpu().numpy(), ge) and \
                np.array_equal(rp["pos_tgroup"][q, :len(pos)].cpu().numpy(), tg) and int(rp["ngroups"][q]) == ngr
        # AP variants against the restated reference functions
        from oracle import reference_metrics as RM
        ppos = list(pos) + ([n] if drop else [])
        if len(ppos):
            ok = ok and RM.compute_ap(ppos, len(ppos)) == st["ap_trapz"][q]
            pp = np.asarray(ppos) + 1
            for t, k in enumerate(kap):
                kq = min(max(pp), k)
                ok = ok and (pp <= kq).sum() / kq == st["prs"][q, t]
        else:
            ok = ok and np.isnan(st["ap_trapz"][q])
        ps = 0.0
        for c, rnk in enumerate(pos):
            ps += (c + 1) / (rnk + 1)
        ok = ok and ps == st["prec_sum"][q] and st["first"][q] == (pos[0] + 1 if len(pos) else 0)
        ok = ok and all(st["hits_at"][q, t] == (pos < k).sum() for t, k in enumerate(kap))
        if ties:
            keep = ~dropped
            rows = np.flatnonzero(keep)
            order = rows[RO.order_row(s[q][rows], largest)]
            if r.any():
                want = RM.average_precision_ranked(s[q][order], r[order])
                ok = ok and want == aps[q]
            else:
                ok = ok and np.isnan(aps[q])
        bad += 0 if ok else 1
    print(f"nq={nq} ng={ng} mode={mode} ties={ties} drop={drop} quant={quant} largest={largest}: {'OK' if bad == 0 else f'{bad} BAD rows'}", flush=True)
    return bad

if __name__ == "__main__":
    bad = 0
    for ng in (1, 2, 37, 400, 2000, 5000, 40000):
        for ties in (False, True):
            bad += one(5, ng, 0, ties, True)
    bad += one(6, 3000, 1, True, False)
    bad += one(6, 3000, 1, False, True, largest=False)
    bad += one(4, 30000, 0, True, True, quant=4)     # massive ties -> refinement
    bad += one(4, 30000, 0, False, True, quant=4)
    bad += one(3, 70000, 0, True, False, quant=1000)
    s = np.zeros((2, 20000), np.float32)              # every score equal
    sd = torch.from_numpy(s).cuda(); lab = torch.zeros(20000, dtype=torch.int64).cuda()
    rp = FR.rank_of_positives(sd, 0, lab[:2], lab, ties=True)
    ok = np.array_equal(rp["pos_ranks"][0].cpu().numpy(), np.arange(20000)) and int(rp["ngroups"][0]) == 1 and \
        bool((rp["pos_ge"][0] == 20000).all())
    print("all-equal row:", "OK" if ok else "BAD"); bad += 0 if ok else 1
    # timing at the NIH scale
    for ng, nq in ((112000, 1024),):
        sd = torch.randn((nq, ng), device="cuda")
        lab = torch.randint(0, 3, (ng,), device="cuda")
        for ties in (False, True):
            FR.rank_of_positives(sd, 0, lab[:nq], lab, drop_self=True, ties=ties); torch.cuda.synchronize()
            t0 = time.perf_counter()
            rp = FR.rank_of_positives(sd, 0, lab[:nq], lab, drop_self=True, ties=ties); torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            print(f"rank_of_positives {nq} x {ng} ties={ties}: {dt*1e3:.2f} ms -> {nq*ng/dt/1e9:.2f} G items/s; 112k rows: {dt*112000/nq:.3f} s")
            t0 = time.perf_counter(); st = FR.ap_from_ranks(rp, (1, 5, 10), True); torch.cuda.synchronize()
            print(f"  ap_from_ranks: {(time.perf_counter()-t0)*1e3:.2f} ms")
            if ties:
                t0 = time.perf_counter(); ap = FR.ap_sklearn_from_ranks(rp); torch.cuda.synchronize()
                print(f"  ap_sklearn_from_ranks: {(time.perf_counter()-t0)*1e3:.2f} ms")
    print("TOTAL BAD", bad)
    sys.exit(1 if bad else 0)
