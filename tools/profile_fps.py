"""FPS on the K8 shapes (16 clouds of 16384 points) for ncu. Development aid."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dv = importlib.import_module("deepvcp-pointcloud-registration_b200")
lib = importlib.import_module("deepvcp-pointcloud-registration_b200._lib")
F_ = dv.functional

B, N = 8, 16384
src, tgt, R, t = dv.synthetic.make_batch("kitti", list(range(B)), N)
both = torch.cat([src, tgt], 0).cuda()
g = torch.Generator().manual_seed(5)
st2 = torch.randint(0, N, (2 * B,), generator=g)
for _ in range(2):
    F_.fps(lib.cloud_cm(both), both.device, both.dtype, 2 * B, N, N, st2, want64=False, want32=True)
torch.cuda.synchronize()
