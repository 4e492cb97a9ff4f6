"""Per-phase cycles of rank_positives_kernel (needs a -DRP_TIMING build selected through KNN_LIB)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from b200knn import fullrank as FR, _lib
nq, ng = int(sys.argv[1]) if len(sys.argv) > 1 else 592, 112000
ties = len(sys.argv) > 2 and sys.argv[2] == "ties"
sd = torch.randn((nq, ng), device="cuda")
lab = torch.randint(0, 3, (ng,), device="cuda")
lib = _lib.load()
FR.rank_of_positives(sd, 0, lab[:nq], lab, drop_self=True, ties=ties)
buf = (ctypes.c_ulonglong * 8)()
lib.knn_rank_timing(buf, 1)
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record(); FR.rank_of_positives(sd, 0, lab[:nq], lab, drop_self=True, ties=ties); t1.record()
lib.knn_rank_timing(buf, 0)
names = ["sample+map", "histogram", "bin scan", "scatter/partition", "refine", "resolve/buckets", "ties/outputs"]
tot = sum(buf[:7])
print(f"{nq} x {ng} ties={ties}: {t0.elapsed_time(t1):.2f} ms (incl. host); per row cycles:")
for n, c in zip(names, buf[:7]):
    print(f"  {n:14s} {c / nq:10.0f} cyc/row  {100.0 * c / tot:5.1f}%")
