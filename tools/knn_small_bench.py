"""KNN variants on the M64 shapes (B=64, N=1024, 64 x 5^3 queries, K=32). Development aid."""
import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dv = importlib.import_module("deepvcp-pointcloud-registration_b200")
F_ = dv.functional
lib = importlib.import_module("deepvcp-pointcloud-registration_b200._lib")
dev = torch.device("cuda")
B, N, G = 64, int(os.environ.get("NPTS", 1024)), 5
src, tgt, R, t = dv.synthetic.make_batch("modelnet", list(range(B)), N)
tgt = tgt.to(dev)
g = torch.Generator().manual_seed(1)
centres = torch.stack([tgt[b, :3, torch.randint(0, N, (64,), generator=g)].T for b in range(B)]).double()
cand = F_.candidates(centres, 0.8, 0.4).view(B, -1, 3)
index = F_.build_index(lib.cloud_cm(tgt), dev, B, N)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(5):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return round(best, 3)
ref = F_.knn(lib.cloud_cm(tgt), dev, B, N, cand, 32, want64=False, want32=True)
out = {"brute": timeit(lambda: F_.knn(lib.cloud_cm(tgt), dev, B, N, cand, 32, want64=False, want32=True))}
for name, ch, zl in (("chain1", 1, 1), ("zline", G, G), ("slab", G * G, G), ("kp", G * G * G, G)):
    fn = lambda: F_.knn_indexed(index, 0, dev, B, N, cand, 32, chain=ch, zline=zl, want64=False, want32=True)
    r = fn()
    assert torch.equal(r[2], ref[2]) and torch.equal(r[0], ref[0])
    out[name] = timeit(fn)
print(out)
