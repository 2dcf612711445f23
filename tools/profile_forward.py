"""One K8 forward + pose solve between cudaProfilerStart/Stop (after a warm-up
forward), for `ncu --profile-from-start off`. Development aid."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dv = importlib.import_module("deepvcp-pointcloud-registration_b200")


def main():
    B = int(os.environ.get("DVCP_B", "8"))
    N = int(os.environ.get("DVCP_N", "16384"))
    dev = torch.device("cuda")
    src, tgt, R, t = dv.synthetic.make_batch("kitti", list(range(B)), N)
    src, tgt, R, t = src.to(dev), tgt.to(dev), R.to(dev), t.view(B, 3, 1).to(dev)
    torch.manual_seed(0)
    model = dv.DeepVCP(use_normal=False, npoint=N, r=2.0, s=0.4).to(dev).eval()
    g = torch.Generator().manual_seed(1000)
    starts = (torch.randint(0, N, (B,), generator=g), torch.randint(0, 64, (B,), generator=g),
              torch.randint(0, N, (B,), generator=g))

    def step():
        kp, vcp = model(src, tgt, R, torch.zeros(1, 3), starts=starts)
        return dv.pose_from_forward(kp, vcp, R, t)

    step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
