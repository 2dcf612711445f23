"""Throughput mode: a stream of batches with the two halves of the forward on two CUDA streams.

One registration forward starts with farthest point sampling, a chain of dependent
selections that occupies few SMs for a long time (DeepVCP.extract_features); everything
after it (KNN, embedding, CPG: DeepVCP.match) fills the GPU. For a stream of batches the
first half of batch k+1 therefore runs on its own stream beside the second half of batch k:

    feature stream   FE(0) FE(1)        FE(2)        FE(3) ...
    match stream           M(0)         M(1)         M(2)  ...

FE(k) is released when M(k - depth) has finished, i.e. (depth 2) exactly when M(k-1)
starts, so the sampling coincides with the KNN of the previous batch and is over before
that batch's persistent embedding / CPG kernels want the whole GPU. Each batch still
runs the complete path (DeepVCP.forward + svd_optimization, deepVCP.py:24-110,
deepVCP_loss.py:57-90) and results come back in submission order; nothing is shared
between batches except the read-only weights. The reference has no counterpart (batch
size 1, one stream, train.py:39,105)."""
import torch

from .deepVCP_loss import pose_from_forward
from .sharding import pack_poses


class StreamedRegistration:
    def __init__(self, model, depth=2):
        dev = model.cpg.conv1.weight.device
        if dev.type != "cuda":
            raise RuntimeError("StreamedRegistration needs the model on a CUDA device")
        if depth < 1:
            raise ValueError("depth >= 1")
        self.model, self.dev, self.depth = model, dev, depth
        self.fe_stream = torch.cuda.Stream(device=dev)
        self.match_stream = torch.cuda.Stream(device=dev)
        self.streams = [self.fe_stream, self.match_stream]
        self.done = []         # completion event of every batch submitted since the last collect()
        self.pending = []      # (done event, poses) in submission order
        self.t_init = torch.zeros(1, 3)
        self.small_sampling_ctas = False   # sampling with half-size CTAs (measured: no gain on B200 at K8)
        self.timing = False    # development: completion events carry timestamps
        self.trace = []

    def submit(self, src, tgt, R_init, R_true, t_true, starts=None, host_out=None, t_init=None):
        """Enqueue one batch: src, tgt [B,C_in,N], R_init / R_true [B,3,3], t_true [B,3,1] (host or
        device tensors). host_out: optional pinned [B,12] float64 tensor the poses are copied into."""
        k = len(self.done)
        # The host may run at most depth + 1 batches ahead of the device: beyond that the caching allocator
        # cannot yet reuse the blocks of finished batches (their stream-use events have not completed), every
        # batch would get freshly cudaMalloc'ed memory, and the pool would grow with the queue length.
        if k > self.depth:
            self.done[k - self.depth - 1].synchronize()
        cur = torch.cuda.current_stream(self.dev)
        to = lambda x: x.to(self.dev, non_blocking=True)
        if self.depth == 1:
            # strictly one batch at a time: both halves on the match stream
            fs = self.match_stream
        else:
            fs = self.fe_stream
        fs.wait_stream(cur)                                   # inputs produced on the caller's stream
        with torch.cuda.stream(fs):
            if self.depth > 1 and k >= self.depth:
                fs.wait_event(self.done[k - self.depth])      # run ahead by at most `depth` batches
            fe = self.model.extract_features(to(src), to(tgt), starts, concurrent=self.small_sampling_ctas)
            Ri, Rt, tt = to(R_init), to(R_true), to(t_true)
            ev_fe = torch.cuda.Event(enable_timing=self.timing)
            ev_fe.record(fs)
        ms = self.match_stream
        with torch.cuda.stream(ms):
            ms.wait_event(ev_fe)
            kp, vcp = self.model.match(fe, Ri, t_init=t_init)   # t_init is only read in intended mode (Q6)
            R2, t2 = pose_from_forward(kp, vcp, Rt, tt, quirks=self.model.quirks)
            poses = pack_poses(R2, t2)
            if host_out is not None:
                host_out.copy_(poses, non_blocking=True)
            ev = torch.cuda.Event(enable_timing=self.timing)
            ev.record(ms)
        self.done.append(ev)
        if self.timing:
            self.trace.append((ev_fe, ev))
        # The first half's tensors were allocated on the feature stream and consumed on the match stream:
        # tell the caching allocator, then let go of them. (Holding them until collect() would make every
        # batch allocate fresh memory, and cudaMalloc synchronises the device.)
        if fs is not ms:
            for t in list(fe.values()) + [Ri, Rt, tt] + [getattr(fe["index"], n, None) for n in ("sorted_pt", "bucket_box")]:
                if torch.is_tensor(t) and t.is_cuda:
                    t.record_stream(ms)
        self.pending.append((ev, poses))
        return k

    def collect(self):
        """Wait for everything submitted so far; returns the [B,12] pose tensors in submission order."""
        out = []
        for ev, poses in self.pending:
            ev.synchronize()
            out.append(poses)
        self.pending = []
        self.done = []
        return out
