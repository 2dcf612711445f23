"""Development probe: e2e (pinned host in, pinned host out, wall clock) of the eager and graph-replay pipelines."""
import importlib, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dv = importlib.import_module("deepvcp-pointcloud-registration_b200")
dev = torch.device("cuda")
B, N = 8, 16384
src, tgt, R, t = dv.synthetic.make_batch("kitti", list(range(B)), N)
torch.manual_seed(0)
model = dv.DeepVCP(use_normal=False, npoint=N, r=2.0, s=0.4).to(dev).eval()
g = torch.Generator().manual_seed(1000)
starts = (torch.randint(0, N, (B,), generator=g), torch.randint(0, 64, (B,), generator=g), torch.randint(0, N, (B,), generator=g))
h = [x.pin_memory() for x in (src, tgt, R, t.view(B, 3, 1))]
d = [x.to(dev) for x in h]
hp = [torch.empty(B, 12, dtype=torch.float64).pin_memory() for _ in range(64)]


def run(pipe, inputs, steps=40, host_out=True, label=""):
    for i in range(8):
        pipe.submit(inputs[0], inputs[1], inputs[2], inputs[2], inputs[3], starts=starts, host_out=hp[i] if host_out else None)
    pipe.collect()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ts = []
    for i in range(steps):
        a = time.perf_counter()
        pipe.submit(inputs[0], inputs[1], inputs[2], inputs[2], inputs[3], starts=starts, host_out=hp[i] if host_out else None)
        ts.append(time.perf_counter() - a)
    pipe.collect()
    dt = time.perf_counter() - t0
    print("%-34s %7.1f pairs/s  %.3f ms/step  host submit median %.3f ms max %.3f" % (label, B * steps / dt, dt / steps * 1e3, sorted(ts)[len(ts) // 2] * 1e3, max(ts) * 1e3))


eager = dv.StreamedRegistration(model, depth=2)
graph = dv.GraphedRegistration(model, B, 3, N, depth=2)
for rep in range(2):
    run(eager, d, label="eager, device in")
    run(eager, h, label="eager, pinned host in/out")
    run(graph, d, label="graphs, device in")
    run(graph, h, label="graphs, pinned host in/out")
    run(graph, h, host_out=False, label="graphs, pinned host in, dev out")
