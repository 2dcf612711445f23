"""A few launches of the CPG kernel at the K8 shape (for ncu captures)."""
import importlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dv = importlib.import_module("deepvcp-pointcloud-registration_b200")
F_ = dv.functional
dev = torch.device("cuda")
M, G = 512, 11
C = G ** 3
g = torch.Generator().manual_seed(0)
net = dv.cpg().to(dev)
src = torch.randn(M, 32, generator=g).to(dev)
tgt = torch.randn(M, 32 * C, generator=g).to(dev)
cand = torch.randn(M, C, 3, generator=g).to(dev)
path = {"tc": F_.CPG_TC, "tcz": F_.CPG_TCZ, "fused": F_.CPG_FUSED}[sys.argv[1] if len(sys.argv) > 1 else "tc"]
for _ in range(3):
    F_.cpg(src, tgt, 0, cand, G, net.params(), path=path)
torch.cuda.synchronize()
