"""Per-phase cycle counts of the tensor-core CPG kernel (library built with DVCP_NVCC_EXTRA=-DDVCP_CPG_TIMING);
CTA 0, thread 0, summed over the volumes that CTA processes."""
import ctypes, importlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dv = importlib.import_module("deepvcp-pointcloud-registration_b200")
lib = importlib.import_module("deepvcp-pointcloud-registration_b200._lib")
F_ = dv.functional
dev = torch.device("cuda")
L = lib.lib()
L.dvcp_debug_cpg_timing.argtypes = [ctypes.c_void_p]
names = ["volume prologue", "fill octet", "MMA issue + wait", "TMEM -> smem", "conv2", "conv3", "softmax + vcp", "loop barrier", "zero A planes (tcz)", "lo half: stage + barrier (tcz)", "drain (tcz)"]
for M, G, path in ((512, 11, "TC"), (512, 11, "TCZ"), (4096, 5, "TC"), (4096, 5, "TCZ"), (512, 7, "TC"), (512, 7, "TCZ"),
                   (4096, 6, "TC"), (4096, 6, "TCZ")):
    C = G ** 3
    g = torch.Generator().manual_seed(0)
    net = dv.cpg().to(dev)
    src = torch.randn(M, 32, generator=g).to(dev)
    tgt = torch.randn(M, 32 * C, generator=g).to(dev)
    cand = torch.randn(M, C, 3, generator=g).to(dev)
    pid = getattr(F_, "CPG_" + path)
    for _ in range(2):
        F_.cpg(src, tgt, 0, cand, G, net.params(), path=pid)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        F_.cpg(src, tgt, 0, cand, G, net.params(), path=pid)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    buf = (ctypes.c_longlong * 16)()
    assert L.dvcp_debug_cpg_timing(buf) == 0
    tot = sum(buf[:11])
    print("%s M = %d, G = %d: %.4f ms per call (instrumented build); %d cycles in CTA 0" % (path, M, G, ms, tot))
    for n, v in zip(names, buf):
        print("  %-24s %10d %5.1f%%" % (n, v, 100.0 * v / tot))
