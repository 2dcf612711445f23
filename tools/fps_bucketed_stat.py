"""Step statistics and time of the one-CTA-per-cloud sampling kernel's wide rounds on a K8 batch (statistics need a
library built with DVCP_NVCC_EXTRA=-DDVCP_FPS_TIMING)."""
import ctypes, importlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dv = importlib.import_module("deepvcp-pointcloud-registration_b200")
lib = importlib.import_module("deepvcp-pointcloud-registration_b200._lib")
F = dv.functional
src, tgt, R, t = dv.synthetic.make_batch("kitti", list(range(8)), 16384)
both = torch.cat([src, tgt], 0).cuda()
idx = F.build_index(lib.cloud_cm(both), both.device, 16, 16384)
st = torch.randint(0, 16384, (16,), generator=torch.Generator().manual_seed(1))
run = lambda: F.fps_indexed(lib.cloud_cm(both), both.device, 16, 16384, 16384, st, idx, concurrent=2)
L = lib.lib()
stat = getattr(L, "dvcp_debug_fps_bucketed_stat", None) if hasattr(L, "dvcp_debug_fps_bucketed_stat") else None
buf = (ctypes.c_longlong * 8)()
run()
torch.cuda.synchronize()
if stat is not None:
    stat.argtypes = [ctypes.c_void_p, ctypes.c_int]
    stat(buf, 1)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
run()
b.record()
torch.cuda.synchronize()
print("kernel %.3f ms" % a.elapsed_time(b))
if stat is not None:
    stat(buf, 1)
    steps, n, A = buf[0], buf[1], buf[2]
    print("steps %d, candidates per step %.1f, accepted per step %.1f" % (steps, n / steps, A / steps))
    for name, v in zip(["update (box tests, distances, bucket tops)", "threshold + candidates (2 barriers)", "rank", "pair tests + barrier",
                        "resolution (warp 0) + barrier"], [buf[4], buf[5], buf[6], buf[7], buf[3]]):
        print("  %-44s %8.0f cycles per step" % (name, v / steps))
