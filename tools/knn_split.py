"""Old (index walk) vs pooled KNN kernel on the dense half and the sparse half of the K8 key-point groups."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import knn_bench as kb
F_, dev = kb.F_, kb.dev
B, N, G = 8, 16384, 11
C = G ** 3
cand, tg, index = kb.workload("kitti", B, N, G, False)
ng = B * 64
stats = torch.zeros(8 + 4 * ng, dtype=torch.int64, device=dev)
F_.knn_groups(index, 0, dev, B, N, cand, 32, C, G, 0.4, pool_cap=4096, want64=False, want32=True, stats=stats)
torch.cuda.synchronize()
npool = stats[8 + ng: 8 + 2 * ng].cpu().view(B, 64)
order = npool.argsort(dim=1)            # ascending pool size per batch item
cg = cand.view(B, 64, C, 3)
for name, sel in (("sparse32", order[:, :32]), ("dense32", order[:, 32:]), ("densest16", order[:, 48:]), ("sparsest16", order[:, :16])):
    q = torch.stack([cg[b, sel[b].to(dev)] for b in range(B)]).reshape(B, -1, 3).contiguous()
    t_old = kb.timeit(lambda: F_.knn_indexed(index, 0, dev, B, N, q, 32, chain=G, want64=False, want32=True))
    t_new = kb.timeit(lambda: F_.knn_groups(index, 0, dev, B, N, q, 32, C, G, 0.4, pool_cap=4096, want64=False, want32=True))
    mean_np = float(torch.gather(npool, 1, sel).double().mean())
    print(name, "groups/item", sel.shape[1], "mean npool %.0f" % mean_np, "old ms %.3f" % t_old[1], "pool ms %.3f" % t_new[1])
