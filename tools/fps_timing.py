"""Per-phase cycle counts of the cluster FPS kernel (library built with -DDVCP_FPS_TIMING)."""
import ctypes, importlib, os, sys
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
L = lib.lib()
buf = (ctypes.c_longlong * 24)()
L.dvcp_debug_fps_timing.argtypes = [ctypes.c_void_p]
print(L.dvcp_debug_fps_timing(buf))
names = ["(1) tests", "bar", "(2) apply", "bar", "(3) local cand", "bar", "(3b)+(4) push", "cluster.sync", "(5) gather+bar",
         "rank sort", "bar", "pair tests", "bar", "greedy(w0)", "bar", "loop"]
tot = sum(buf[:16])
for n, v in zip(names, buf):
    print("%-18s %12d %5.1f%%" % (n, v, 100.0 * v / tot))
print("total cycles", tot, "steps", buf[16], "sum n", buf[17], "sum A", buf[18], "sum local cnt", buf[19], "n>CAP", buf[20])

sm = (ctypes.c_int * 1024)()
L.dvcp_debug_fps_smid.argtypes = [ctypes.c_void_p]
L.dvcp_debug_fps_smid(sm)
ids = list(sm)[:16 * 8]
print("SMs used by the 128 sampling CTAs:", len(set(ids)), "distinct; per cluster:", [sorted(ids[c * 8:c * 8 + 8]) for c in range(4)])
