"""Throughput mode: consecutive batches of pairs in flight on alternating CUDA streams.

One registration forward starts with farthest point sampling, a chain of dependent
selections that occupies few SMs for a long time; everything after it (KNN, embedding,
CPG) fills the GPU. A stream of batches therefore overlaps the sampling of batch i + 1
with the dense stages of batch i. Each batch still runs the complete path
(DeepVCP.forward + svd_optimization, deepVCP.py:24-110, deepVCP_loss.py:57-90) and
results are returned in submission order; nothing is shared between batches except the
read-only weights. The reference has no counterpart (batch size 1, one stream,
train.py:39,105)."""
import torch

from .deepVCP_loss import pose_from_forward
from .sharding import pack_poses


class StreamedRegistration:
    def __init__(self, model, depth=2):
        dev = model.cpg.conv1.weight.device
        if dev.type != "cuda":
            raise RuntimeError("StreamedRegistration needs the model on a CUDA device")
        self.model, self.dev, self.depth = model, dev, depth
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(depth)]
        self.pending = []      # (done event, poses tensor) in submission order
        self.n = 0
        self.t_init = torch.zeros(1, 3)

    def submit(self, src, tgt, R_init, R_true, t_true, starts=None, host_out=None):
        """Enqueue one batch: src, tgt [B,C_in,N], R_init / R_true [B,3,3], t_true [B,3,1] (host or
        device tensors). host_out: optional pinned [B,12] float64 tensor the poses are copied into."""
        s = self.streams[self.n % self.depth]
        self.n += 1
        s.wait_stream(torch.cuda.current_stream(self.dev))   # inputs produced on the caller's stream
        with torch.cuda.stream(s):
            to = lambda x: x.to(self.dev, non_blocking=True)
            src_d, tgt_d, Ri, Rt, tt = to(src), to(tgt), to(R_init), to(R_true), to(t_true)
            kp, vcp = self.model(src_d, tgt_d, Ri, self.t_init, starts=starts)
            R2, t2 = pose_from_forward(kp, vcp, Rt, tt)
            poses = pack_poses(R2, t2)
            if host_out is not None:
                host_out.copy_(poses, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(s)
        self.pending.append((ev, poses))
        return len(self.pending) - 1

    def collect(self):
        """Wait for everything submitted so far; returns the [B,12] pose tensors in submission order."""
        out = []
        for ev, poses in self.pending:
            ev.synchronize()
            out.append(poses)
        self.pending = []
        return out
