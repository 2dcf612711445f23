"""Time of one training step (train-mode forward + deepVCP_loss + backward + Adam) at the reference's own operating
point (ModelNet-shaped pair, N = 10000 with normals, 6^3 candidates, B = 1: train.py:39,93-125) and at a KITTI-shaped
pair (N = 16384, 11^3). Development aid; the reference's CPU forward alone takes ~48 s at the KITTI shape."""
import importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dv = importlib.import_module("deepvcp-pointcloud-registration_b200")
dev = torch.device("cuda")
for kind, N, r, fused in (("modelnet", 10000, 1.0, True), ("modelnet", 10000, 1.0, False), ("kitti", 16384, 2.0, True)):
    src, tgt, R, t = dv.synthetic.make_batch(kind, [0], N)
    torch.manual_seed(0)
    model = dv.DeepVCP(use_normal=kind == "modelnet", npoint=N, r=r, s=0.4).to(dev).train()
    optim = torch.optim.Adam(model.parameters(), lr=1e-3)
    src, tgt, R, t = src.to(dev), tgt.to(dev), R.to(dev), t.view(1, 3, 1).to(dev)
    times = []
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.time()
        kp, vcp = dv.training.forward(model, src, tgt, R, torch.zeros(1, 3), fused_embedding=fused)
        optim.zero_grad()
        loss, Rp, tp = dv.deepVCP_loss(kp, vcp, R, t, alpha=0.5)
        loss.backward(); optim.step()
        torch.cuda.synchronize(); times.append(time.time() - t0)
    print(kind, N, "fused_embedding" if fused else "autograd_embedding", "step ms", [round(x * 1e3, 1) for x in times],
          "peak GB", round(torch.cuda.max_memory_allocated() / 1e9, 2), "loss", float(loss.detach()))
    torch.cuda.reset_peak_memory_stats()
