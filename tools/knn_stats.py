"""Pool statistics of the line-pool KNN kernel on the K8 shapes. Needs a library built with
DVCP_NVCC_EXTRA=-DDVCP_KNN_STATS (development only)."""
import ctypes
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dv = importlib.import_module("deepvcp-pointcloud-registration_b200")
F_ = dv.functional
lib = importlib.import_module("deepvcp-pointcloud-registration_b200._lib")

dev = torch.device("cuda")
B, N, G = 8, 16384, 11
src, tgt, R, t = dv.synthetic.make_batch("kitti", list(range(B)), N)
src, tgt, R = src.to(dev), tgt.to(dev), R.to(dev)
torch.manual_seed(0)
model = dv.DeepVCP(use_normal=False, npoint=N, r=2.0, s=0.4).to(dev).eval()
g = torch.Generator().manual_seed(5)
starts = (torch.randint(0, N, (B,), generator=g), torch.randint(0, 64, (B,), generator=g),
          torch.randint(0, N, (B,), generator=g))
model(src, tgt, R, torch.zeros(1, 3), starts=starts, keep_stages=True)
cand = model.last["candidates"].view(B, -1, 3)
index = F_.build_index(lib.cloud_cm(tgt), dev, B, N)
L = lib.lib()
fn = L.dvcp_knn_debug_stats
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
out = (ctypes.c_ulonglong * 8)()
fn(None, 1)
F_.knn_indexed(index, 0, dev, B, N, cand, 32, chain=G * G, zline=G, want64=False, want32=True)
fn(out, 1)
names = ["lines pooled", "lines overflowed", "sum npool", "max npool", "query fallbacks", "sum cnt", "first lines pooled",
         "first lines overflowed"]
d = dict(zip(names, [int(v) for v in out]))
lines = d["lines pooled"] + d["lines overflowed"] + d["first lines pooled"] + d["first lines overflowed"]
d["mean npool"] = d["sum npool"] / max(lines, 1)
d["mean cnt"] = d["sum cnt"] / max(d["lines pooled"] * G, 1)
print(d)
