"""Per-group cycles of the pooled KNN kernel (needs a build with -DDVCP_KNN_GROUP_STATS)."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import knn_bench as kb
cand, tg, index = kb.workload("kitti", 8, 16384, 11, False)
cap = int(sys.argv[1]) if len(sys.argv) > 1 else 0
ng = 8 * 64
stats = torch.zeros(8 + 4 * ng, dtype=torch.int64, device=kb.dev)
kb.F_.knn_groups(index, 0, kb.dev, 8, 16384, cand, 32, 1331, 11, 0.4, pool_cap=cap, want64=False, want32=True)
stats.zero_()
kb.F_.knn_groups(index, 0, kb.dev, 8, 16384, cand, 32, 1331, 11, 0.4, pool_cap=cap, want64=False, want32=True, stats=stats)
torch.cuda.synchronize()
st = stats.cpu()
cyc, npool, probe, pool = (st[8 + i * ng: 8 + (i + 1) * ng].double() for i in range(4))
order = cyc.argsort(descending=True)
print("total Mcycles %.1f  max %.2f  mean %.3f  sum/444 %.3f" % (cyc.sum() / 1e6, cyc.max() / 1e6, cyc.mean() / 1e6, cyc.sum() / 444e6))
print("probe share %.3f  pool-build share %.3f" % (probe.sum() / cyc.sum(), (pool - probe).sum() / cyc.sum()))
for j in order[:25].tolist() + order[-5:].tolist():
    print("group %3d cycles %8.0fk npool %5d probe %6.0fk build %6.0fk" % (j, cyc[j] / 1e3, npool[j], probe[j] / 1e3, (pool[j] - probe[j]) / 1e3))
import numpy as np
c = np.corrcoef(npool.numpy(), cyc.numpy())[0, 1]
print("corr(npool, cycles) = %.3f" % c)
