"""One launch of the pooled KNN kernel on the K8 workload (for ncu captures)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import knn_bench as kb
cand, tg, index = kb.workload("kitti", 8, 16384, 11, False)
cap = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for _ in range(2):
    kb.F_.knn_groups(index, 0, kb.dev, 8, 16384, cand, 32, 1331, 11, 0.4, pool_cap=cap, want64=False, want32=True)
torch.cuda.synchronize()
