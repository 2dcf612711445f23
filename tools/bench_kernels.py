"""Per-kernel timings on the K8 shapes (CUDA events, L2 flushed between runs).
Development aid; bench.py is the contract."""
import argparse
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dv = importlib.import_module("deepvcp-pointcloud-registration_b200")
F_ = dv.functional
lib = importlib.import_module("deepvcp-pointcloud-registration_b200._lib")


def timeit(fn, flush, iters=5):
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
    return min(ts), sum(ts) / len(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--N", type=int, default=16384)
    ap.add_argument("--G", type=int, default=11)
    ap.add_argument("--kind", default="kitti")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    dev = torch.device("cuda")
    B, N, G = args.B, args.N, args.G
    r = dv.synthetic.grid_radius(G)
    src, tgt, R, t = dv.synthetic.make_batch(args.kind, list(range(B)), N)
    src, tgt, R = src.to(dev), tgt.to(dev), R.to(dev)
    C_in = src.shape[1]
    torch.manual_seed(0)
    model = dv.DeepVCP(use_normal=C_in == 6, npoint=N, r=r, s=0.4).to(dev).eval()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator().manual_seed(5)
    starts = (torch.randint(0, N, (B,), generator=g), torch.randint(0, 64, (B,), generator=g),
              torch.randint(0, N, (B,), generator=g))
    model(src, tgt, R, torch.zeros(1, 3), starts=starts, keep_stages=True)
    L = model.last
    both = torch.cat([src, tgt], 0)
    st2 = torch.cat([starts[0], starts[2]])
    fps2 = torch.cat([L["src_fps"], L["tgt_fps"]], 0).contiguous()
    sa = model.FE1.sa1
    D = C_in - 3
    cand = L["candidates"].view(B, -1, 3)
    kd, _, ki = F_.knn(lib.cloud_cm(tgt), dev, B, N, cand, 32, want64=False, want32=True)
    tfeat = L["tgt_fe_feat"].contiguous()
    tgt_dfe = L["tgt_dfe"].contiguous()
    K = 64
    Cc = G ** 3
    res = {}
    index = F_.build_index(lib.cloud_cm(both), dev, 2 * B, N) if F_.SpatialIndex.indexable(N) else None
    runs = {
        "fps": lambda: F_.fps(lib.cloud_cm(both), dev, both.dtype, 2 * B, N, N, st2, want64=False, want32=True),
        "sa_layer": lambda: F_.sa_layer(lib.cloud_cm(both), lib.cloud_cm(both[:, 3:]) if D else None, D, fps2, 2 * B,
                                        N, N, sa.radius, sa.nsample, sa.folded(), dev, want_xyz=False, index=index),
        "build_index": lambda: F_.build_index(lib.cloud_cm(both), dev, 2 * B, N),
        "knn": lambda: F_.knn(lib.cloud_cm(tgt), dev, B, N, cand, 32, want64=False, want32=True),
        "knn_indexed": lambda: F_.knn_indexed(index, B, dev, B, N, cand, 32, chain=G * G, zline=G, want64=False, want32=True),
        "knn_indexed_zline": lambda: F_.knn_indexed(index, B, dev, B, N, cand, 32, chain=G, want64=False, want32=True),
        "knn_chain1": lambda: F_.knn_indexed(index, B, dev, B, N, cand, 32, chain=1, want64=False, want32=True),
        "knn_chain2": lambda: F_.knn_indexed(index, B, dev, B, N, cand, 32, chain=2, want64=False, want32=True),
        "knn_chain4": lambda: F_.knn_indexed(index, B, dev, B, N, cand, 32, chain=4, want64=False, want32=True),
        "knn_2lines": lambda: F_.knn_indexed(index, B, dev, B, N, cand, 32, chain=2 * G, zline=G, want64=False, want32=True),
        "knn_3lines": lambda: F_.knn_indexed(index, B, dev, B, N, cand, 32, chain=3 * G, zline=G, want64=False, want32=True),
        "knn_4lines": lambda: F_.knn_indexed(index, B, dev, B, N, cand, 32, chain=4 * G, zline=G, want64=False, want32=True),
        "knn_6lines": lambda: F_.knn_indexed(index, B, dev, B, N, cand, 32, chain=6 * G, zline=G, want64=False, want32=True),
        "knn_indexed_kp": lambda: F_.knn_indexed(index, B, dev, B, N, cand, 32, chain=G * G * G, zline=G, want64=False, want32=True),
        "dfe": lambda: F_.dfe_tgt_fused(cand, lib.cloud_cm(tgt), tfeat, kd, ki, B, N, model.DFE.params(),
                                        lib.QUIRKS_REFERENCE),
        "dfe_tc": lambda: F_.dfe_tgt_tc(cand, lib.cloud_cm(tgt), tfeat, kd, ki, B, N, *model.DFE.tc_operand(),
                                        lib.QUIRKS_REFERENCE),
        "cpg": lambda: F_.cpg(L["src_dfe"].view(B * K, 32), tgt_dfe.view(B * K, Cc * 32), 1,
                              cand.view(B * K, Cc, 3), G, model.cpg.params()),
        "forward": lambda: model(src, tgt, R, torch.zeros(1, 3), starts=starts),
    }
    for name, fn in runs.items():
        if args.only and name not in args.only.split(","):
            continue
        res[name] = [round(v, 4) for v in timeit(fn, flush)]
    print(json.dumps({"B": B, "N": N, "G": G, "kind": args.kind, "env_fps_warps": os.environ.get("DVCP_FPS_WARPS"),
                      "ms_min_avg": res}))


if __name__ == "__main__":
    main()
