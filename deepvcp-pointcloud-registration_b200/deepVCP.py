"""DeepVCP -- deepVCP.py:16-110 of the reference: the registration forward.

Same constructor, forward signature, tensor layouts and state_dict keys; the
hyper-parameters the reference hard-codes become keyword arguments whose
defaults are the reference's literals. forward() is a fixed sequence of
sm_100a kernels (no tensor of the reference larger than the KNN result is
materialised):

    FPS(src), FPS(tgt)                      pointnet2_utils.py:63-84
    SA layer(src), SA layer(tgt)            :110-138,176-202 fused
    weighting MLP + top-K                   weighting_layer.py:26-33
    key-point stage                         deepVCP.py:44-67,86-91,101
    candidate grid                          voxelize.py:19-83
    KNN                                     get_cat_feat_tgt.py:45-52
    gather + normalise + DFE + max          get_cat_feat_tgt.py:53-96, deep_feat_embedding.py:46-60
    CPG                                     cpg.py:27-60

B > 1 means B independent pairs (the reference only runs B = 1).
Reference-mode quirks are on by default and individually
switchable through `quirks` (bits QUIRK_* of _lib.py; QUIRKS_INTENDED = 0 selects
the semantics the reference's code intends for Q3, Q4, Q6, Q7 -- SURVEY 8f rank 2).
"""
import torch
import torch.nn as nn

from . import functional as F_
from ._lib import QUIRK_COST_VOLUME_RESHAPE, QUIRK_FPS_ORDER_MISMATCH, QUIRKS_REFERENCE, cloud_cm, cloud_pm, require_cuda
from .cpg import cpg
from .deep_feat_embedding import feat_embedding_layer
from .deep_feat_extraction import feat_extraction_layer
from .weighting_layer import weighting_layer


class DeepVCP(nn.Module):
    def __init__(self, use_normal, npoint=10000, fe_radius=0.1, fe_nsample=256, K_topk=64, nsample=32,
                 r=1.0, s=0.4, group_radius=1, quirks=QUIRKS_REFERENCE, chained_fe=False):
        super().__init__()
        # chained_fe: the three-layer feature extraction the reference's file intends;
        # default = what the reference can run (sa1 only)
        self.FE1 = feat_extraction_layer(use_normal=use_normal, npoint=npoint, radius=fe_radius,
                                         nsample=fe_nsample, chained=chained_fe)
        self.WL = weighting_layer()
        self.DFE = feat_embedding_layer()
        self.cpg = cpg()
        self.K_topk, self.nsample, self.r, self.s = K_topk, nsample, r, s
        self.group_radius = group_radius
        self.quirks = quirks
        self.last = None   # stage tensors of the most recent forward (for tests / inspection)
        self.dfe_tensor_cores = True    # tcgen05 embedding (collapsed affine map, 3xTF32); False = FP32 CUDA-core kernel
        # KNN through per-key-point shared-memory pools (knn_pool.cu, dvcp_knn_groups): exact, but measured SLOWER
        # than the index walk per query on B200 (K8: 3.8 ms against 1.65 ms) -- kept as an option
        self.knn_pools_min_n = 0
        self.knn_pools = False
        self.profile = False   # record a CUDA event after every stage (bench.py reads them)
        self._events = None

    def _side_stream(self, dev):
        if getattr(self, "_side", None) is None or self._side.device != dev:
            self._side = torch.cuda.Stream(device=dev, priority=-1)
        return self._side

    def _identity(self, B, N, dev):
        t = getattr(self, "_ident", None)
        if t is None or t.shape != (B, N) or t.device != dev:
            self._ident = torch.arange(N, dtype=torch.int32, device=dev).repeat(B, 1).contiguous()
        return self._ident

    def draw_starts(self, B, N):
        """The three FPS start draws of one forward in the reference's order:
        FE(src) -> key-point grouping -> FE(tgt) (deepVCP.py:29,54,72)."""
        if self.FE1.chained:   # one draw per set-abstraction layer
            draw3 = lambda: torch.stack([F_.draw_fps_start(B, N) for _ in range(3)])
            a = draw3()
            k = F_.draw_fps_start(B, self.K_topk)
            return (a, k, draw3())
        return (F_.draw_fps_start(B, N), F_.draw_fps_start(B, self.K_topk), F_.draw_fps_start(B, N))

    def stage_times_ms(self):
        """Per-stage device time of the last profiled forward (after a synchronize)."""
        ev = self._events or []
        return {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(ev, ev[1:])}

    def forward(self, src_pts, tgt_pts, R_init, t_init, starts=None, keep_stages=False, topk_override=None):
        """src_pts, tgt_pts [B,C_in,N]; R_init [B,3,3] float64; t_init [1,3] (unused by
        the reference, quirk Q6) -> (src_keypts [B,K,3], tgt_vcp [B,K,3])."""
        if self.training:
            # train.py:93-105: batch statistics in the BatchNorm layers and an autograd graph (training.py)
            from . import training
            return training.forward(self, src_pts, tgt_pts, R_init, t_init, starts=starts, keep_stages=keep_stages,
                                    topk_override=topk_override)
        fe = self.extract_features(src_pts, tgt_pts, starts)
        return self.match(fe, R_init, keep_stages=keep_stages, topk_override=topk_override, t_init=t_init)

    def _mark(self, name, dev):
        if self._events is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(torch.cuda.current_stream(dev))
            self._events.append((name, e))

    def prepare_index(self, src_pts, tgt_pts):
        """The part of extract_features() that does not depend on the FPS starts: the two clouds side by side and
        their spatial index (16 single-CTA sorts). A stream of batches builds it one batch further ahead, off the
        critical path (pipeline.GraphedRegistration). Returns None when the forward takes another route."""
        B, C_in, N = src_pts.shape
        if (self.FE1.chained or src_pts.dtype != torch.float32 or tgt_pts.dtype != torch.float32 or
                not (F_.SpatialIndex.indexable(N) and N > 2048 and 2 * self.FE1.sa1.npoint >= N)):
            return None
        with torch.no_grad():
            both = torch.cat([src_pts, tgt_pts], dim=0)
            return dict(both=both, index=F_.build_index(cloud_cm(both), both.device, 2 * B, N))

    def extract_features(self, src_pts, tgt_pts, starts=None, concurrent=False, prepared=None):
        """First half of forward(): feature extraction of both clouds (deepVCP.py:29,72). Returns the
        state match() continues from. Split out so that a stream of batches can run this half (few SMs,
        long) beside the second half of the previous batch (pipeline.StreamedRegistration). prepared: the
        result of prepare_index() on the same clouds."""
        if self.training:
            raise RuntimeError("extract_features / match are the inference kernels: call .eval() first (train mode goes "
                               "through forward(), see training.py)")
        for x in (src_pts, tgt_pts):
            if x.dtype not in (torch.float32, torch.float64):
                raise RuntimeError("clouds must be float32 or float64, got %s" % x.dtype)
        dev = self.cpg.conv1.weight.device
        if dev.type != "cuda":
            raise RuntimeError("DeepVCP (b200) needs its parameters on a CUDA device; there is no CPU fallback")
        B, C_in, N = src_pts.shape
        if starts is None:
            starts = self.draw_starts(B, N)
        # host tensors (pinned or not) are accepted at the boundary and copied once
        src = src_pts.to(dev, non_blocking=True).contiguous()
        tgt = tgt_pts.to(dev, non_blocking=True).contiguous()
        require_cuda(src, tgt)
        S = self.FE1.sa1.npoint
        sa = self.FE1.sa1
        mlp = sa.folded()
        D = C_in - 3
        self._events = [] if self.profile else None
        mark = lambda name: self._mark(name, dev)
        if src.dtype == torch.float64 or tgt.dtype == torch.float64:
            return self._extract_features_mixed(src, tgt, starts, B, N, C_in, S, D, dev)

        with torch.no_grad():
            mark("begin")
            # feature extraction: source and target clouds go through each kernel in ONE
            # launch (2B independent clouds); features come out in FPS order
            both = prepared["both"] if prepared else torch.cat([src, tgt], dim=0)
            if self.FE1.chained:
                # three chained set-abstraction layers + fc through the module (one launch sequence for 2B clouds)
                s0, s2 = torch.as_tensor(starts[0]), torch.as_tensor(starts[2])
                st2 = torch.cat([s0.reshape(-1, B), s2.reshape(-1, B)], dim=1)      # [3, 2B] or [1, 2B]
                _, feat2, fps2 = self.FE1(both, start=st2 if st2.shape[0] == 3 else st2[0], return_fps=True)
                mark("fps")
                mark("sa_layer")
                index = F_.build_index(cloud_cm(both), dev, 2 * B, N) if F_.SpatialIndex.indexable(N) else None
                return dict(src=src, tgt=tgt, both=both, index=index, fps2=fps2, feat2=feat2.contiguous(),
                            starts=starts, B=B, N=N, C_in=C_in, dev=dev)
            st2 = torch.cat([torch.as_tensor(starts[0]).reshape(-1), torch.as_tensor(starts[2]).reshape(-1)])
            feat_cloud = cloud_cm(both[:, 3:, :]) if D else None
            if F_.SpatialIndex.indexable(N) and N > 2048 and 2 * S >= N:
                # FPS is a chain of dependent selections that leaves most of the GPU idle, and the SA
                # features of a point do not depend on its FPS rank: build the index, then run the SA
                # layer over the points in their ORIGINAL order beside the sampling, and put the rows
                # into FPS order afterwards (feat_fps[s] = feat_orig[fps[s]]).
                index = prepared["index"] if prepared else F_.build_index(cloud_cm(both), dev, 2 * B, N)
                main = torch.cuda.current_stream(dev)
                ev_index = torch.cuda.Event()
                ev_index.record(main)
                # the sampling goes to a HIGH-priority stream and is enqueued first: its clusters of 8
                # co-scheduled CTAs must get their SMs before the SA layer's many small CTAs fill the GPU
                hp = self._side_stream(dev)
                with torch.cuda.stream(hp):
                    hp.wait_event(ev_index)
                    _, fps2 = F_.fps_indexed(cloud_cm(both), dev, 2 * B, N, S, st2, index, concurrent=concurrent)
                    ev_fps = torch.cuda.Event()
                    ev_fps.record(hp)
                feat_orig = F_.sa_layer_all(cloud_cm(both), feat_cloud, D, self._identity(2 * B, N, dev), 2 * B, N,
                                            sa.radius, sa.nsample, mlp, dev, index)
                main.wait_event(ev_fps)
                fps2.record_stream(main)
                mark("fps")
                feat2 = F_.gather_rows(feat_orig, fps2)
                mark("sa_layer")
            else:
                index = F_.SpatialIndex(2 * B, N, dev) if F_.SpatialIndex.indexable(N) else None
                _, fps2 = F_.fps(cloud_cm(both), dev, both.dtype, 2 * B, N, S, st2, want64=False, want32=True,
                                 index=index)
                mark("fps")
                _, feat2 = F_.sa_layer(cloud_cm(both), feat_cloud, D, fps2, 2 * B, N, S, sa.radius, sa.nsample, mlp,
                                       dev, want_xyz=False, index=index)
                mark("sa_layer")
        return dict(src=src, tgt=tgt, both=both, index=index, fps2=fps2, feat2=feat2, starts=starts, B=B, N=N,
                    C_in=C_in, dev=dev)

    def _extract_features_mixed(self, src, tgt, starts, B, N, C_in, S, D, dev):
        """Feature extraction when a cloud is float64 -- what the reference's loaders hand over
        (ModelNet40Dataset.py:38,92: both float64; KITTIDataset.py:84,97: float32 scan, float64 target).
        torch's promotion rules are followed cloud by cloud: FPS distances and the ball query in double
        (pointnet2_utils.py:80-82,35-40,100-102), relative coordinates formed in double and cast to float
        before the shared MLP (:198). A float32 cloud goes through the float32 kernels. The target is
        additionally kept as float32 for the KNN (knn_cuda casts its inputs) and the embedding."""
        if self.FE1.chained:
            raise NotImplementedError("chained_fe with float64 clouds")
        sa = self.FE1.sa1
        mlp = sa.folded()
        mark = lambda name: self._mark(name, dev)
        with torch.no_grad():
            mark("begin")
            fps, feat = [], []
            for pts, st in ((src, starts[0]), (tgt, starts[2])):
                xyz_cloud = cloud_cm(pts)
                feat_cloud = cloud_cm(pts[:, 3:, :]) if D else None
                if pts.dtype == torch.float64:
                    _, f32i = F_.fps(xyz_cloud, dev, pts.dtype, B, N, S, st, want64=False, want32=True)
                    _, ft = F_.sa_layer_f64(xyz_cloud, feat_cloud, D, f32i, B, N, S, sa.radius, sa.nsample, mlp, dev,
                                            want_xyz=False)
                else:
                    idx = F_.SpatialIndex(B, N, dev) if F_.SpatialIndex.indexable(N) else None
                    _, f32i = F_.fps(xyz_cloud, dev, pts.dtype, B, N, S, st, want64=False, want32=True, index=idx)
                    _, ft = F_.sa_layer(xyz_cloud, feat_cloud, D, f32i, B, N, S, sa.radius, sa.nsample, mlp, dev,
                                        want_xyz=False, index=idx)
                fps.append(f32i)
                feat.append(ft)
            mark("fps")
            mark("sa_layer")
            tgt32 = tgt.float()                       # knn_cuda: ref.float()  (get_cat_feat_tgt.py:45,52)
            index = F_.build_index(cloud_cm(tgt32), dev, B, N) if F_.SpatialIndex.indexable(N) else None
        return dict(src=src, tgt=tgt32, both=None, index=index, index_lo=0, fps2=torch.cat(fps), feat2=torch.cat(feat),
                    starts=starts, B=B, N=N, C_in=C_in, dev=dev)

    def match(self, fe, R_init, keep_stages=False, topk_override=None, t_init=None):
        """Second half of forward(): key-point selection, candidates, KNN, embedding, CPG
        (deepVCP.py:33-110) on the state extract_features() returned."""
        st = self.match_search(fe, R_init, keep_stages=keep_stages, topk_override=topk_override, t_init=t_init)
        return self.match_finish(st, keep_stages=keep_stages)

    def match_search(self, fe, R_init, keep_stages=False, topk_override=None, t_init=None):
        """match() up to and including the KNN; match_finish() continues with the tensor-core kernels."""
        return self.match_knn(self.match_select(fe, R_init, keep_stages=keep_stages, topk_override=topk_override,
                                                t_init=t_init), keep_stages=keep_stages)

    def match_select(self, fe, R_init, keep_stages=False, topk_override=None, t_init=None):
        """The small kernels at the head of match(): key-point selection, key-point stage, candidate lattice
        (deepVCP.py:33-67,86-91,101). A handful of short launches that leave the GPU idle: the depth >= 3 pipeline
        runs them at the end of the FEATURE half, beside the dense kernels of an earlier batch."""
        src, tgt, index, fps2, feat2, starts = fe["src"], fe["tgt"], fe["index"], fe["fps2"], fe["feat2"], fe["starts"]
        B, N, dev = fe["B"], fe["N"], fe["dev"]
        ilo = fe.get("index_lo", B)   # target clouds are batch items ilo..ilo+B-1 of the index
        K, ns = self.K_topk, self.nsample
        R = R_init.to(dev, non_blocking=True)
        require_cuda(R)
        mark = lambda name: self._mark(name, dev)
        with torch.no_grad():
            sfps, tfps = fps2[:B], fps2[B:]
            sfeat, tfeat = feat2[:B], feat2[B:]
            # key-point selection
            scores = self.WL.scores(sfeat)
            topk = F_.topk(scores, K) if topk_override is None else topk_override.to(dev).view(B, K)
            mark("weighting_topk")
            dfe = self.DFE.params()
            kp_rows, kp_feat = topk, sfeat
            if not (self.quirks & QUIRK_FPS_ORDER_MISMATCH):
                # intended wiring (Q5): row k of the feature table belongs to point fps[k], so the key-point
                # coordinates are those of fps[topk] (the reference gathers src_pts[:, topk], deepVCP.py:35,46), the
                # key-points' neighbourhood features are the key-points' own rows (the reference reads rows 0..63 of
                # the whole table with key-point-local indices, :61), and the target features are addressed by
                # original point index like the KNN result that indexes them (get_cat_feat_tgt.py:85)
                kp_rows = torch.gather(sfps.long(), 1, topk)
                kp_feat = F_.gather_rows(sfeat, topk.int())
                t_orig = torch.empty_like(tfeat)
                t_orig.scatter_(1, tfps.long().unsqueeze(-1).expand(-1, -1, tfeat.shape[-1]), tfeat)
                tfeat = t_orig
            keypts, picked, cat, src_dfe, centres = F_.keypoint_stage(
                src, kp_rows, starts[1], kp_feat, R, self.group_radius, ns, dfe, self.quirks,
                want_cat=keep_stages, want_picked=keep_stages, t_init=t_init)
            # candidates, KNN, target-side embedding
            G = F_.grid_size(self.r, self.s)
            cand = F_.candidates(centres, self.r, self.s, G)                 # [B,K,C,3]
            mark("keypoint_candidates")
            C = G * G * G
            # point-major float4 copy of the target xyz for the tensor-core embedding: one 16-byte load per gathered neighbour
            tgt4 = F_.pack_xyz4(cloud_cm(tgt), dev, B, N) if self.dfe_tensor_cores else None
        return dict(fe=fe, tfeat=tfeat, sfps=sfps, tfps=tfps, sfeat=sfeat, scores=scores, topk=topk, keypts=keypts,
                    picked=picked, cat=cat, src_dfe=src_dfe, centres=centres, cand=cand, G=G, C=C, dfe=dfe, tgt4=tgt4)

    def match_knn(self, st, keep_stages=False):
        """The K nearest target points of every candidate (get_cat_feat_tgt.py:45,52) on the state of match_select()."""
        fe = st["fe"]
        tgt, index, B, N, dev = fe["tgt"], fe["index"], fe["B"], fe["N"], fe["dev"]
        ilo = fe.get("index_lo", B)   # target clouds are batch items ilo..ilo+B-1 of the index
        K, ns, cand, C, G = self.K_topk, self.nsample, st["cand"], st["C"], st["G"]
        mark = lambda name: self._mark(name, dev)
        with torch.no_grad():
            if index is not None and self.knn_pools and N >= self.knn_pools_min_n:
                # one CTA per key-point: the target points around its candidate lattice are pooled in shared memory
                kd, ki64, ki32 = F_.knn_groups(index, ilo, dev, B, N, cand.view(B, K * C, 3), ns, group=C, zline=G,
                                               cell=self.s, want64=keep_stages, want32=True)
            elif index is not None:
                kd, ki64, ki32 = F_.knn_indexed(index, ilo, dev, B, N, cand.view(B, K * C, 3), ns, chain=G,
                                                want64=keep_stages, want32=True)
            elif F_.SpatialIndex.knn_indexable(N):
                # clouds above the sampling kernels' index capacity: multi-CTA index of the targets for the KNN
                tindex = F_.build_index(cloud_cm(tgt), dev, B, N, big=True)
                kd, ki64, ki32 = F_.knn_indexed(tindex, 0, dev, B, N, cand.view(B, K * C, 3), ns, chain=G,
                                                want64=keep_stages, want32=True)
            else:
                kd, ki64, ki32 = F_.knn(cloud_cm(tgt), dev, B, N, cand.view(B, K * C, 3), ns, want64=keep_stages,
                                        want32=True)
            mark("knn")
        st = dict(st)
        st.update(kd=kd, ki64=ki64, ki32=ki32)
        return st

    def match_finish(self, st, keep_stages=False):
        """Embedding, CPG (deepVCP.py:93-110) on the state match_search() returned."""
        fe = st["fe"]
        tgt, B, N, dev = fe["tgt"], fe["B"], fe["N"], fe["dev"]
        K = self.K_topk
        tfeat, sfps, tfps, sfeat, scores, topk = st["tfeat"], st["sfps"], st["tfps"], st["sfeat"], st["scores"], st["topk"]
        keypts, picked, cat, src_dfe, centres, cand = st["keypts"], st["picked"], st["cat"], st["src_dfe"], st["centres"], st["cand"]
        kd, ki64, ki32, G, C, dfe = st["kd"], st["ki64"], st["ki32"], st["G"], st["C"], st["dfe"]
        mark = lambda name: self._mark(name, dev)
        with torch.no_grad():
            reshape_quirk = bool(self.quirks & QUIRK_COST_VOLUME_RESHAPE)
            fm = False   # tgt_dfe held feature-major per key-point ([B,K,32,C])?
            if self.dfe_tensor_cores:
                b_hi, b_lo = self.DFE.tc_operand()
                tgt4 = st["tgt4"]
                # Reference mode: cpg.py:34 re-reads the LOGICAL [32, C] order of the permuted tensor (Q4); the
                # embedding kernel writes exactly that order, so CPG finds a voxel's 32 values contiguous
                # (layout 0) and its tensor-core kernel applies. Intended mode wants [C, 32] as it is: also layout 0.
                fm = reshape_quirk
                tgt_dfe = F_.dfe_tgt_tc(cand.view(B, K * C, 3), cloud_pm(tgt4), tfeat, kd, ki32, B, N, b_hi, b_lo,
                                        self.quirks, feature_major_c=C if fm else 0)   # [B,K*C,32] or [B,K,32,C]
            else:
                tgt_dfe = F_.dfe_tgt_fused(cand.view(B, K * C, 3), cloud_cm(tgt), tfeat, kd, ki32, B, N, dfe,
                                           self.quirks)
            mark("dfe")
            # corresponding point generation. layout 0: the flat tensor IS the logical [32, C] order the reference
            # reshapes (or, in intended mode, cost[c, f] = (src[f] - tgt[c, f])^2 read as it lies); layout 1: the
            # flat tensor is [C, 32] and the kernel applies the reference's permute + reshape itself
            vcp, logits = F_.cpg(src_dfe.view(B * K, 32), tgt_dfe.reshape(B * K, C * 32),
                                 1 if (reshape_quirk and not fm) else 0,
                                 cand.view(B * K, C, 3), G, self.cpg.params(), want_logits=keep_stages)
            if fm:
                tgt_dfe = tgt_dfe.permute(0, 1, 3, 2)        # logical [B,K,C,32] view for inspection
            mark("cpg")
        if keep_stages:
            self.last = dict(src_fps=sfps, tgt_fps=tfps, src_fe_feat=sfeat, tgt_fe_feat=tfeat, scores=scores,
                             topk_idx=topk, src_keypts_full=keypts, picked_idx=picked, src_cat=cat,
                             src_dfe=src_dfe, centres=centres, candidates=cand, knn_dist=kd, knn_idx=ki64,
                             tgt_dfe=tgt_dfe.reshape(B, K, C, 32), logits=logits, vcp=vcp.view(B, K, 3))
        return keypts[:, :, :3], vcp.view(B, K, 3)
