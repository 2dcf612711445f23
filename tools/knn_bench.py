"""KNN stage of the K8 / M64 workloads in isolation: the index-walk kernel (knn.cu) against the pooled
kernel (knn_pool.cu) for several pool capacities, with the pooled kernel's profiling counters.
Development / evidence tool (bench.py is the contract). Prints one JSON object."""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dv = importlib.import_module("deepvcp-pointcloud-registration_b200")
lib = importlib.import_module("deepvcp-pointcloud-registration_b200._lib")
F_ = dv.functional
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sorted(ts)[len(ts) // 2]


def workload(kind, B, N, G, use_normal):
    r = dv.synthetic.grid_radius(G)
    src, tgt, R, t = dv.synthetic.make_batch(kind, list(range(B)), N)
    torch.manual_seed(0)
    model = dv.DeepVCP(use_normal=use_normal, npoint=N, r=r, s=0.4).to(dev).eval()
    g = torch.Generator().manual_seed(1000)
    starts = (torch.randint(0, N, (B,), generator=g), torch.randint(0, 64, (B,), generator=g),
              torch.randint(0, N, (B,), generator=g))
    model(src.to(dev), tgt.to(dev), R.to(dev), torch.zeros(1, 3), starts=starts, keep_stages=True)
    cand = model.last["candidates"].view(B, -1, 3).contiguous()
    tg = tgt.to(dev)
    index = F_.build_index(lib.cloud_cm(tg), dev, B, N)
    return cand, tg, index


def main():
    out = {}
    for name, kind, B, N, G, un in (("K8", "kitti", 8, 16384, 11, False), ("M64", "modelnet", 64, 1024, 5, True)):
        cand, tg, index = workload(kind, B, N, G, un)
        C = G ** 3
        res = {}
        d0, _, i0 = F_.knn_indexed(index, 0, dev, B, N, cand, 32, chain=G, want64=False, want32=True)
        res["indexed_ms(min,median)"] = timeit(lambda: F_.knn_indexed(index, 0, dev, B, N, cand, 32, chain=G, want64=False, want32=True))
        for cap in ([0, 4096] if name == "K8" else [0]):
            stats = torch.zeros(8, dtype=torch.int64, device=dev)
            d, _, i = F_.knn_groups(index, 0, dev, B, N, cand, 32, C, G, 0.4, pool_cap=cap, want64=False, want32=True, stats=stats)
            ok = bool(torch.equal(i, i0) and torch.equal(d, d0))
            ms = timeit(lambda: F_.knn_groups(index, 0, dev, B, N, cand, 32, C, G, 0.4, pool_cap=cap, want64=False, want32=True))
            st = stats.cpu().tolist()
            nq = B * cand.shape[1]
            res["pool_%d" % cap] = {"ms(min,median)": ms, "equal_to_indexed": ok, "certified": st[0] / nq, "uncertified": st[1] / nq,
                                    "overflow": st[2] / nq, "cold": st[3] / nq, "no_pool": st[4] / nq,
                                    "pool_points_per_group": st[5] / (B * 64), "admitted_per_query": st[6] / nq,
                                    "scanned_per_query": st[7] / nq}
        out[name] = res
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
