"""BASELINE.json configs[1] (M64) and configs[4] (isolated kernel sweep) on one B200:
ModelNet-shaped batch throughput, KNN / ball query over N = 4k..128k, candidate grids 5^3..21^3,
batched Kabsch over 1e5 pairs. Prints one JSON object; every kernel is spot-checked against the
CPU oracle on a small sample. Development / evidence tool (bench.py is the contract)."""
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
from oracle import stages  # noqa: E402  (checker only)

dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(iters):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def kitti_cloud(n, seed):
    g = torch.Generator().manual_seed(seed)
    rho = (torch.randn(n, generator=g) * 25.0).abs().clamp(max=80.0)
    az = torch.rand(n, generator=g) * 6.2831853
    u, v = torch.rand(n, generator=g), torch.rand(n, generator=g)
    z = torch.where(u < 0.7, v * 3.0 - 2.0, v * 5.0 + 1.0)
    p = torch.stack([rho * torch.cos(az), rho * torch.sin(az), z], 1)
    return (torch.round(p * 10.0) / 10.0).float().unsqueeze(0)          # [1, n, 3], 0.1 m lattice (ties)


def main():
    out = {}
    # ---- configs[1]: ModelNet-shaped batch B = 64, N = 1024, 64 key-points, 5^3, K = 32 ----
    B, N = 64, 1024
    src, tgt, R, t = dv.synthetic.make_batch("modelnet", list(range(B)), N)
    src, tgt, R, t = src.to(dev), tgt.to(dev), R.to(dev), t.view(B, 3, 1).to(dev)
    torch.manual_seed(0)
    model = dv.DeepVCP(use_normal=True, npoint=N, r=0.8, s=0.4).to(dev).eval()
    g = torch.Generator().manual_seed(7)
    starts = (torch.randint(0, N, (B,), generator=g), torch.randint(0, 64, (B,), generator=g),
              torch.randint(0, N, (B,), generator=g))

    def step():
        kp, vcp = model(src, tgt, R, torch.zeros(1, 3), starts=starts)
        return dv.pose_from_forward(kp, vcp, R, t)
    ms = timeit(step, 5)
    out["M64_modelnet_B64_N1024_5cubed"] = {"ms_per_batch": round(ms, 3), "pairs_per_s": round(B / ms * 1e3, 1)}

    # ---- configs[4]: KNN and ball query over N ----
    knn, ball = {}, {}
    for n in (4096, 8192, 16384, 32768, 65536, 131072):
        cloud = kitti_cloud(n, n).to(dev)
        gq = torch.Generator().manual_seed(1)
        centres = cloud[0, torch.randint(0, n, (64,), generator=gq)].double().unsqueeze(0)
        cand = F_.candidates(centres, 2.0, 0.4).view(1, -1, 3)                 # 64 x 11^3 queries
        index = F_.build_index(lib.cloud_pm(cloud), dev, 1, n, big=True)
        fn = lambda: F_.knn_indexed(index, 0, dev, 1, n, cand, 32, chain=11, want64=False, want32=True)
        kind = "indexed" if n <= 16384 else "indexed (multi-CTA index build)"
        ms = timeit(fn)
        ms_index = timeit(lambda: F_.build_index(lib.cloud_pm(cloud), dev, 1, n, big=True))
        ms_brute = timeit(lambda: F_.knn(lib.cloud_pm(cloud), dev, 1, n, cand, 32, want64=False, want32=True))
        d, _, i32 = fn()
        sel = torch.arange(0, cand.shape[1], 997)
        dref, iref = stages.knn(cloud.cpu(), cand[:, sel].cpu(), 32)
        ok = bool(torch.equal(i32[:, sel].cpu().long(), iref) and torch.equal(d[:, sel].cpu(), dref))
        Q = cand.shape[1]
        knn[str(n)] = {"ms": round(ms, 3), "index_build_ms": round(ms_index, 3), "brute_force_ms": round(ms_brute, 3),
                       "kernel": kind, "queries": Q, "bit_exact_sample": ok,
                       "algorithmic_gbs": round((12 * n + 12 * Q + 8 * Q * 32) / ms / 1e6, 1)}
        q = cloud[:, :4096].contiguous()
        msb = timeit(lambda: F_.ball_query(1.0, 32, cloud, q))
        bq = F_.ball_query(1.0, 32, cloud, q)
        bref = stages.query_ball_point(1.0, 32, cloud.cpu(), q[:, :64].cpu())
        ball[str(n)] = {"ms": round(msb, 3), "queries": 4096, "bit_exact_sample": bool(torch.equal(bq[:, :64].cpu(), bref))}
    out["knn_k32_64x11cubed_queries_vs_N"] = knn
    out["ball_query_r1_ns32_4096_queries_vs_N"] = ball

    # ---- configs[4]: candidate grids 5^3 .. 21^3 (candidates + KNN + embedding + cpg of one pair, N = 16384) ----
    grids = {}
    n = 16384
    src, tgt, R, t = dv.synthetic.make_batch("kitti", [0], n)
    src, tgt, R = src.to(dev), tgt.to(dev), R.to(dev)
    for G in (5, 7, 11, 15, 21):
        r = dv.synthetic.grid_radius(G)
        torch.manual_seed(0)
        m = dv.DeepVCP(use_normal=False, npoint=n, r=r, s=0.4).to(dev).eval()
        st = (torch.tensor([1]), torch.tensor([2]), torch.tensor([3]))
        m.profile = True
        ms = timeit(lambda: m(src, tgt, R, torch.zeros(1, 3), starts=st))
        torch.cuda.synchronize()
        tm = m.stage_times_ms()
        grids["%d^3" % G] = {"forward_ms": round(ms, 3), "knn_ms": round(tm["knn"], 3), "dfe_ms": round(tm["dfe"], 3),
                             "cpg_ms": round(tm["cpg"], 3), "cpg_kernel": "fused" if G <= 11 else "per-layer"}
    out["candidate_grid_sweep_B1_N16384"] = grids

    # ---- configs[4]: batched Kabsch, 1e5 pairs x 64 points ----
    g = torch.Generator().manual_seed(99)
    Bk = 100000
    x = torch.randn(Bk, 3, 64, generator=g, dtype=torch.float64)
    ang = torch.rand(Bk, generator=g, dtype=torch.float64) * 6.28
    Rz = torch.zeros(Bk, 3, 3, dtype=torch.float64)
    Rz[:, 0, 0], Rz[:, 0, 1], Rz[:, 1, 0], Rz[:, 1, 1], Rz[:, 2, 2] = ang.cos(), -ang.sin(), ang.sin(), ang.cos(), 1.0
    tt = torch.randn(Bk, 3, 1, generator=g, dtype=torch.float64)
    y = Rz @ x + tt + 0.01 * torch.randn(Bk, 3, 64, generator=g, dtype=torch.float64)
    kab = {}
    for name, xx, yy in (("f64", x, y), ("f32", x.float(), y.float())):
        xd, yd = xx.to(dev), yy.to(dev)
        ms = timeit(lambda: F_.kabsch(xd, yd))
        Rg, tg = F_.kabsch(xd, yd)
        err = float((Rg.cpu() - Rz).abs().max())
        kab[name] = {"ms": round(ms, 3), "pairs_per_s": round(Bk / ms * 1e3), "max_abs_R_err_vs_truth": round(err, 5),
                     "algorithmic_gbs": round(Bk * (2 * 64 * 3 * xx.element_size() + 96) / ms / 1e6, 1)}
    out["kabsch_1e5_pairs_x64"] = kab
    print(json.dumps(out))


if __name__ == "__main__":
    main()
