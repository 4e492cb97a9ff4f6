import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from b200knn import fullrank as FR
nq, ng = int(sys.argv[1]) if len(sys.argv) > 1 else 592, 112000
ties = len(sys.argv) > 2 and sys.argv[2] == "ties"
sd = torch.randn((nq, ng), device="cuda")
lab = torch.randint(0, 3, (ng,), device="cuda")
for _ in range(2):
    rp = FR.rank_of_positives(sd, 0, lab[:nq], lab, drop_self=True, ties=ties)
torch.cuda.synchronize()
if ties:
    FR.ap_sklearn_from_ranks(rp)
torch.cuda.synchronize()
