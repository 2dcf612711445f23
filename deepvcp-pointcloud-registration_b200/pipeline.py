"""Throughput mode: a stream of batches with the two halves of the forward on two CUDA streams.

One registration forward starts with farthest point sampling, a chain of dependent
selections that occupies few SMs for a long time (DeepVCP.extract_features); everything
after it (KNN, embedding, CPG: DeepVCP.match) fills the GPU. For a stream of batches the
first half of batch k+1 therefore runs on its own stream beside the second half of batch k:

    feature stream   FE(0) FE(1)        FE(2)        FE(3) ...
    match stream           M(0)         M(1)         M(2)  ...

FE(k) is released when M(k - depth) has finished, i.e. (depth 2) exactly when M(k-1)
starts, so the sampling coincides with the KNN of the previous batch and is over before
that batch's embedding / CPG kernels want the whole GPU.

depth >= 3 (what bench.py uses) trades latency for throughput: the sampling runs as ONE CTA per
cloud (dvcp_fps_indexed, concurrent = 2: 3.8 ms instead of 2.3 ms for the 16 clouds of a K8 batch, but
16 SMs instead of 122, under half the SM time), and the feature halves of depth - 1 batches are in
flight at once on their own streams beside the match half of an earlier batch:

    feature stream 0   FE(0)......  FE(2)......  FE(4)......
    feature stream 1         FE(1)......  FE(3)......
    match stream                   M(0)  M(1)  M(2)  M(3) ...

The sampling CTAs hold their SMs for milliseconds, which is why the embedding and CPG kernels are
launched as several CTAs per SM slot (common.cuh, DVCP_DFE_WAVES): whichever SMs are free take the
work; the short launches at the head of the match half (DeepVCP.match_select) run at the end of the
feature half. K8 on one B200: 3.69 -> 3.10 ms per batch. Each batch still
runs the complete path (DeepVCP.forward + svd_optimization, deepVCP.py:24-110,
deepVCP_loss.py:57-90) and results come back in submission order; nothing is shared
between batches except the read-only weights. The reference has no counterpart (batch
size 1, one stream, train.py:39,105)."""
import torch

from .deepVCP_loss import pose_from_forward
from .sharding import pack_poses


def auto_depth(N):
    """The pipeline depth that gives the highest throughput on one B200: large clouds (the sampling of a batch takes
    milliseconds) gain from depth 3 (one sampling CTA per cloud, two feature halves in flight: K8 3.74 -> 3.10 ms per
    batch); for small clouds the sampling is short and the plain two-stream overlap is best (M64: 2.26 ms at depth 2,
    2.34 ms at depth 3)."""
    return 3 if N > 2048 else 2


def _pipeline_shape(depth, fe_streams, sampling):
    """Defaults of the two pipeline classes: depth <= 2: one feature stream, cluster sampling (mode 0);
    depth >= 3: depth - 1 feature streams, one sampling CTA per cloud (mode 2)."""
    if depth < 1:
        raise ValueError("depth >= 1")
    if fe_streams is None:
        fe_streams = max(1, depth - 1)
    if sampling is None:
        sampling = 2 if depth >= 3 else 0
    if sampling not in (0, 1, 2) or fe_streams < 1:
        raise ValueError("sampling in (0, 1, 2), fe_streams >= 1")
    return fe_streams, sampling


class StreamedRegistration:
    def __init__(self, model, depth=2, fe_streams=None, sampling=None):
        """depth: batches in flight (1 = one at a time). fe_streams: feature halves in flight at once (default
        depth - 1). sampling: the `concurrent` mode of dvcp_fps_indexed (default 0 = clusters up to depth 2,
        2 = one CTA per cloud from depth 3 on)."""
        dev = model.cpg.conv1.weight.device
        if dev.type != "cuda":
            raise RuntimeError("StreamedRegistration needs the model on a CUDA device")
        n_fe, self.sampling = _pipeline_shape(depth, fe_streams, sampling)
        self.model, self.dev, self.depth = model, dev, depth
        self.fe_stream = torch.cuda.Stream(device=dev)
        self.fe_streams = [self.fe_stream] + [torch.cuda.Stream(device=dev) for _ in range(n_fe - 1)]
        self.match_stream = torch.cuda.Stream(device=dev)
        self.streams = self.fe_streams + [self.match_stream]
        self.done = []         # completion event of every batch submitted since the last collect()
        self.pending = []      # (done event, poses) in submission order
        self.t_init = torch.zeros(1, 3)
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
            fs = self.fe_streams[k % len(self.fe_streams)]
        fs.wait_stream(cur)                                   # inputs produced on the caller's stream
        with torch.cuda.stream(fs):
            if self.depth > 1 and k >= self.depth:
                fs.wait_event(self.done[k - self.depth])      # run ahead by at most `depth` batches
            fe = self.model.extract_features(to(src), to(tgt), starts, concurrent=self.sampling)
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


class GraphedRegistration:
    """StreamedRegistration with the two halves of the forward captured ONCE into CUDA graphs and replayed.

    One batch of a fixed shape (B, C_in, N, dtype float32) costs two graph launches instead of ~20 kernel
    launches, a dozen allocator calls and as many ctypes transitions: the host leaves the critical path, which
    is what the max-over-ranks time of a multi-GPU run is sensitive to (8 processes on shared cores). `slots`
    sets of static buffers and graphs are used round-robin, so that batch k+1's feature half runs beside batch
    k's match half exactly as in StreamedRegistration; results are identical to it (same kernels, same order).
    The reference has no counterpart (train.py:105 calls the eager model once per pair)."""

    def __init__(self, model, B, C_in, N, depth=None, fe_streams=None, sampling=None):
        """depth / fe_streams / sampling: as for StreamedRegistration; depth None = auto_depth(N)."""
        dev = model.cpg.conv1.weight.device
        if dev.type != "cuda":
            raise RuntimeError("GraphedRegistration needs the model on a CUDA device")
        from . import functional as F_
        self.model, self.dev, self.depth = model, dev, max(1, depth if depth is not None else auto_depth(N))
        self.n_fe, self.sampling = _pipeline_shape(self.depth, fe_streams, sampling)
        self.B, self.C_in, self.N = B, C_in, N
        import os
        self.ahead_index = os.environ.get("DVCP_AHEAD_INDEX", "1") == "1"    # spatial index built one batch ahead
        self.select_early = self.depth >= 3 and os.environ.get("DVCP_SELECT_EARLY", "1") == "1"
        self.fe_stream = torch.cuda.Stream(device=dev)
        self.fe_streams = [self.fe_stream] + [torch.cuda.Stream(device=dev) for _ in range(self.n_fe - 1)]
        self.match_stream = torch.cuda.Stream(device=dev) if self.depth > 1 else self.fe_stream
        self.copy_stream = torch.cuda.Stream(device=dev)   # input copies run ahead of the (in-order) feature stream
        self.streams = self.fe_streams + [self.match_stream]
        self.slots = []
        self.pending = []
        self.launches_per_batch = 0
        K = model.K_topk
        g = torch.Generator().manual_seed(0)
        for _ in range(self.depth + 1 if self.depth > 1 else 1):
            s = dict(src=torch.zeros(B, C_in, N, device=dev), tgt=torch.zeros(B, C_in, N, device=dev),
                     Ri=torch.eye(3, dtype=torch.float64, device=dev).repeat(B, 1, 1),
                     Rt=torch.eye(3, dtype=torch.float64, device=dev).repeat(B, 1, 1),
                     tt=torch.zeros(B, 3, 1, dtype=torch.float64, device=dev),
                     st=[torch.zeros(B, dtype=torch.int64, device=dev) for _ in range(3)],
                     st_pin=torch.zeros(3, B, dtype=torch.int64).pin_memory(), st_ev=None,
                     free=None)
            # plausible contents for the warm-up / capture runs (the kernels must not see degenerate clouds)
            s["src"].copy_(torch.rand(B, C_in, N, generator=g) * 20 - 10)
            s["tgt"].copy_(torch.rand(B, C_in, N, generator=g) * 20 - 10)
            self.slots.append(s)
        for s in self.slots:
            self._capture(s, F_)

    def _run_index(self, s):
        return self.model.prepare_index(s["src"], s["tgt"])

    def _run_fe(self, s, prepared=None):
        fe = self.model.extract_features(s["src"], s["tgt"], (s["st"][0], s["st"][1], s["st"][2]),
                                         concurrent=self.sampling, prepared=prepared)
        # depth >= 3: the short launches at the head of the match half (key-point selection, key-point stage, candidate
        # lattice) leave the GPU almost idle; at the end of the feature half they run beside an earlier batch's dense kernels
        return self.model.match_select(fe, s["Ri"]) if self.select_early else fe

    def _run_match(self, s, fe):
        if self.select_early:
            kp, vcp = self.model.match_finish(self.model.match_knn(fe))
        else:
            kp, vcp = self.model.match(fe, s["Ri"])
        R2, t2 = pose_from_forward(kp, vcp, s["Rt"], s["tt"], quirks=self.model.quirks)
        return pack_poses(R2, t2)

    def _capture(self, s, F_):
        fs, ms = self.fe_stream, self.match_stream
        cur = torch.cuda.current_stream(self.dev)
        fs.wait_stream(cur)
        with torch.cuda.stream(fs):           # eager warm-up: lazy initialisation must not happen under capture
            for _ in range(2):
                fe = self._run_fe(s, self._run_index(s))
                self._run_match(s, fe)
        fs.synchronize()
        n0 = F_.LAUNCHES
        # The spatial index of a batch (one single-CTA sort per cloud, 0.19 ms at K8) depends on the clouds only:
        # its own graph, replayed on the copy stream right behind the input copies, i.e. as soon as the slot is
        # free -- a whole period before the sampling needs it, beside the previous batches' kernels
        s["prepared"] = None
        if self.ahead_index and self._run_index(s) is not None:
            s["g_index"] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(s["g_index"], stream=fs):
                s["prepared"] = self._run_index(s)
        s["g_fe"] = torch.cuda.CUDAGraph()
        with torch.cuda.graph(s["g_fe"], stream=fs):
            s["fe"] = self._run_fe(s, s["prepared"])
        ms.wait_stream(fs)
        s["g_match"] = torch.cuda.CUDAGraph()
        with torch.cuda.graph(s["g_match"], stream=ms):
            s["poses"] = self._run_match(s, s["fe"])
        self.launches_per_batch = F_.LAUNCHES - n0
        torch.cuda.synchronize(self.dev)

    def submit(self, src, tgt, R_init, R_true, t_true, starts, host_out=None):
        """Enqueue one batch (shapes as given to the constructor; host or device tensors; starts = the three
        FPS start index tensors [B]). Poses are copied into host_out (pinned [B,12] float64) if given."""
        k = len(self.pending)
        s = self.slots[k % len(self.slots)]
        fs, ms, cs = self.fe_streams[k % len(self.fe_streams)], self.match_stream, self.copy_stream
        cur = torch.cuda.current_stream(self.dev)
        cs.wait_stream(cur)
        # The inputs go into the slot's static buffers on their OWN stream: queued behind the previous batch's
        # feature half they would start ~0.2 ms late, the match half of the batch before would take the SMs first
        # and the sampling clusters would have to wait for them (measured: 4.5 -> 5.3 ms per step).
        with torch.cuda.stream(cs):
            if s["free"] is not None:
                cs.wait_event(s["free"])      # the slot's previous batch has left the match half
            s["src"].copy_(src, non_blocking=True)
            s["tgt"].copy_(tgt, non_blocking=True)
            s["Ri"].copy_(R_init, non_blocking=True)
            s["Rt"].copy_(R_true, non_blocking=True)
            s["tt"].copy_(t_true.reshape(self.B, 3, 1), non_blocking=True)
            # FPS starts usually arrive as pageable host tensors: a copy from pageable memory would block the host
            # until the stream gets there (i.e. until the slot is free) -- stage them through the slot's pinned row
            if s["st_ev"] is not None:
                s["st_ev"].synchronize()      # the previous copy out of the pinned row (three batches ago) is done
            for j, (d, v) in enumerate(zip(s["st"], starts)):
                v = torch.as_tensor(v).reshape(-1)
                if v.is_cuda:
                    d.copy_(v, non_blocking=True)
                else:
                    s["st_pin"][j].copy_(v)
                    d.copy_(s["st_pin"][j], non_blocking=True)
            if s["prepared"] is not None:
                s["g_index"].replay()         # needs the clouds only
            s["st_ev"] = torch.cuda.Event()
            s["st_ev"].record(cs)
            for x in (src, tgt, R_init, R_true, t_true) + tuple(v for v in starts if torch.is_tensor(v)):
                if x.is_cuda:
                    x.record_stream(cs)       # the caller's tensor is read on this stream: keep its memory until then
        with torch.cuda.stream(fs):
            fs.wait_event(s["st_ev"])
            if self.depth > 1 and k >= self.depth:
                fs.wait_event(self.pending[k - self.depth][0])   # run ahead by at most `depth` batches
            s["g_fe"].replay()
            ev_fe = torch.cuda.Event()
            ev_fe.record(fs)
        with torch.cuda.stream(ms):
            ms.wait_event(ev_fe)
            s["g_match"].replay()
            out = s["poses"]
            if host_out is not None:
                host_out.copy_(out, non_blocking=True)
            else:
                out = out.clone()            # the static output buffer is overwritten by the slot's next batch
            ev = torch.cuda.Event()
            ev.record(ms)
        s["free"] = ev
        self.pending.append((ev, out if host_out is None else host_out))
        return k

    def collect(self):
        out = []
        for ev, poses in self.pending:
            ev.synchronize()
            out.append(poses)
        self.pending = []
        return out
