"""Per-stage device times of the match half while the next batch's feature half runs beside it
(throughput mode), against the same stages alone. Development aid."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dv = importlib.import_module("deepvcp-pointcloud-registration_b200")
B, N = 8, 16384
dev = torch.device("cuda")
src, tgt, R, t = dv.synthetic.make_batch("kitti", list(range(B)), N)
src, tgt, R, t = src.to(dev), tgt.to(dev), R.to(dev), t.view(B, 3, 1).to(dev)
torch.manual_seed(0)
model = dv.DeepVCP(use_normal=False, npoint=N, r=2.0, s=0.4).to(dev).eval()
g = torch.Generator().manual_seed(1000)
starts = (torch.randint(0, N, (B,), generator=g), torch.randint(0, 64, (B,), generator=g), torch.randint(0, N, (B,), generator=g))
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 2
pipe = dv.StreamedRegistration(model, depth=depth)
for _ in range(4):
    pipe.submit(src, tgt, R, R, t, starts=starts)
pipe.collect()
model.profile = True
evs = []
for _ in range(12):
    pipe.submit(src, tgt, R, R, t, starts=starts)
    evs.append(model._events)     # events of this batch (feature half marks + match half marks)
pipe.collect()
torch.cuda.synchronize()
acc = {}
for ev in evs[2:]:
    for (n0, e0), (n1, e1) in zip(ev, ev[1:]):
        acc.setdefault(n1, []).append(e0.elapsed_time(e1))
print("depth", depth, {k: round(sum(v) / len(v), 3) for k, v in acc.items()})
