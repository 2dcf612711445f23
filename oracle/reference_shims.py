"""Import the UNMODIFIED reference from /root/reference under small shims.

TEST INFRASTRUCTURE, build container only: /root/reference does not exist on
the GPU box. This module is used by tests/golden/make_golden.py to generate
the committed fixtures and by the `needs_reference` CPU tests to re-check the
oracle restatement (oracle/stages.py, oracle/dvcp_oracle.c) against the live
reference. Nothing in the product package imports it.

Shims (none of them edits a reference file):
  S1  stub `matplotlib` / `matplotlib.pyplot`     (voxelize.py:5 imports it)
  S2  stand-in module `knn_cuda` with class KNN   (get_cat_feat_tgt.py:4,
      deepVCP_loss.py:3; the real extension is third-party and absent). The
      stand-in follows the published contract of unlimblue/KNN_CUDA, see
      oracle/dvcp_oracle.c:orc_knn_f32.
  S3  torch.Tensor.cuda -> identity on a CPU box  (get_cat_feat_tgt.py:52)
  S4  feat_extraction_layer.forward := sa1 only   (deep_feat_extraction.py:26-28
      crash at HEAD); sa1.npoint set to N.
  S5  (r, s) override for the literals of deepVCP.py:76-77.
  S6  use_normal=False only: channel count of deepVCP.py:44 taken from the
      input -- done by running the forward on a 6-channel view is
      NOT possible, so S6 re-states lines 24-110 with that one literal changed
      (`forward_s6`), every other line calling the reference's own modules.
"""
import contextlib
import io
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("DVCP_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "deepVCP.py"))


class _KNN:
    """Stand-in for knn_cuda.KNN (S2): float32, direct-form squared distance,
    (distance, index) ascending, sqrt of the kept distances, int64 indices."""

    def __init__(self, k, transpose_mode=False):
        self.k = k
        self.transpose_mode = transpose_mode

    def __call__(self, ref, query):
        from . import native
        with torch.no_grad():
            ref = ref.float()
            query = query.float()
            if not self.transpose_mode:
                ref = ref.transpose(1, 2)
                query = query.transpose(1, 2)
            d, i = native.knn(ref.contiguous(), query.contiguous(), self.k)
            if not self.transpose_mode:
                d = d.transpose(1, 2).contiguous()
                i = i.transpose(1, 2).contiguous()
        return d, i


_mods = None


def load():
    """Returns a namespace of the reference's modules (imported once)."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    # S1
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    # S2
    knn_mod = types.ModuleType("knn_cuda")
    knn_mod.KNN = _KNN
    sys.modules["knn_cuda"] = knn_mod
    # S3
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    names = ["pointnet2_utils", "deep_feat_extraction", "weighting_layer", "voxelize",
             "get_cat_feat_src", "get_cat_feat_tgt", "deep_feat_embedding", "cpg",
             "deepVCP", "deepVCP_loss"]
    ns = types.SimpleNamespace()
    for n in names:
        setattr(ns, n, importlib.import_module(n))
    _mods = ns
    return ns


@contextlib.contextmanager
def quiet():
    """The reference prints inside forward(); swallow it."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def make_model(use_normal: bool, n_points: int, seed: int = 0):
    """DeepVCP(use_normal) with seeded default init, eval mode, S4 applied.

    BatchNorm running statistics are perturbed (seeded) so that the folded
    eval-mode affine is not the identity and parity tests exercise it.
    """
    m = load()
    torch.manual_seed(seed)
    model = m.deepVCP.DeepVCP(use_normal=use_normal)
    g = torch.Generator().manual_seed(seed + 1)
    for mod in model.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.1)
            mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) * 0.5 + 0.75)
            mod.weight.data.copy_(torch.rand(mod.weight.shape, generator=g) * 0.5 + 0.75)
            mod.bias.data.copy_(torch.randn(mod.bias.shape, generator=g) * 0.1)
    model.eval()
    fe = model.FE1
    fe.sa1.npoint = n_points

    def fe_forward(pts, _fe=fe):          # S4
        if _fe.use_normal:
            normal = pts[:, 3:, :]
            xyz = pts[:, :3, :]
        else:
            normal = None
            xyz = pts
        oxyz, opts = _fe.sa1(xyz, normal)
        return oxyz.permute(0, 2, 1), opts.permute(0, 2, 1)

    fe.forward = fe_forward
    return model


def forward(model, src, tgt, R_init, t_init, r: float, s: float, record: dict = None, grad: bool = False):
    """Reference DeepVCP.forward under S5 (and S6 for 3-channel clouds).

    `record`, if given, is filled with the stage-boundary tensors. grad=True keeps the autograd graph
    (training fixtures: train.py:105-123).
    """
    m = load()
    dv = m.deepVCP
    rec = record if record is not None else {}
    orig_vox = dv.voxelize
    orig_cpg_fwd = model.cpg.forward
    orig_gct = dv.Get_Cat_Feat_Tgt
    orig_gcs = dv.Get_Cat_Feat_Src
    orig_sag = dv.sample_and_group
    orig_wl = model.WL.forward
    orig_dfe = model.DFE.forward
    orig_fe = model.FE1.forward

    def vox(pc, _r, _s):                  # S5
        rec["centres"] = pc.detach().clone()
        out = orig_vox(pc, r, s)
        rec["candidates"] = out.detach().clone()
        return out

    def cpg_fwd(a, b, c, _r, _s):         # S5
        rec["src_dfe"] = a.detach().clone()
        rec["tgt_dfe"] = b.detach().contiguous().clone()
        out = orig_cpg_fwd(a, b, c, r, s)
        rec["vcp"] = out.detach().clone()
        return out

    class GCT(orig_gct):
        def forward(self, cand, kp, txyz, tfeat):
            out = super().forward(cand, kp, txyz, tfeat)
            rec["tgt_cat"] = out
            return out

    class GCS(orig_gcs):
        def forward(self, kp, grp, feats):
            out = super().forward(kp, grp, feats)
            rec["src_keypts_full"] = kp.detach().clone()
            rec["src_grouped"] = grp.detach().clone()
            rec["src_keyfeats"] = feats.detach().clone()
            rec["src_cat"] = out.detach().clone()
            return out

    def sag(*a, **k):
        out = orig_sag(*a, **k)
        if k.get("returnidx"):
            rec["kp_new_xyz"] = out[0].detach().clone()
            rec["picked_idx"] = out[2].detach().clone()
        return out

    def wl(X, K=64):
        rec.setdefault("wl_in", X.detach().clone())
        out = orig_wl(X, K)
        rec["topk_idx"] = out.detach().clone()
        return out

    fe_calls = []

    def fe(pts):
        out = orig_fe(pts)
        fe_calls.append((out[0].detach().clone(), out[1].detach().clone()))
        return out

    dv.voxelize = vox
    model.cpg.forward = cpg_fwd
    dv.Get_Cat_Feat_Tgt = GCT
    dv.Get_Cat_Feat_Src = GCS
    dv.sample_and_group = sag
    model.WL.forward = wl
    model.FE1.forward = fe
    try:
        with (contextlib.nullcontext() if grad else torch.no_grad()), quiet():
            if src.shape[1] == 6:
                kp, vcp = model(src, tgt, R_init, t_init)
            else:
                kp, vcp = _forward_s6(model, dv, src, tgt, R_init, t_init)
    finally:
        dv.voxelize = orig_vox
        model.cpg.forward = orig_cpg_fwd
        dv.Get_Cat_Feat_Tgt = orig_gct
        dv.Get_Cat_Feat_Src = orig_gcs
        dv.sample_and_group = orig_sag
        model.WL.forward = orig_wl
        model.FE1.forward = orig_fe
        model.DFE.forward = orig_dfe
    rec["src_fe_xyz"], rec["src_fe_feat"] = fe_calls[0]
    rec["tgt_fe_xyz"], rec["tgt_fe_feat"] = fe_calls[1]
    rec["src_keypts"] = kp.detach().clone()
    return kp, vcp


def _forward_s6(model, dv, src_pts, tgt_pts, R_init, t_init):
    """deepVCP.py:24-110 with the literal 6 of line 44 replaced by the input's
    channel count (S6). Every callee is the reference's own (patched as above)."""
    B = src_pts.shape[0]
    _, src_feat = model.FE1(src_pts)
    idx = model.WL(src_feat)
    idx_u = (idx.unsqueeze(0)).unsqueeze(1).repeat(1, src_pts.shape[1], 1)
    src_keypts = torch.gather(src_pts, 2, idx_u).view(B, 64, src_pts.shape[1])
    _, grouped, picked = dv.sample_and_group(npoint=64, radius=1, nsample=32,
                                             xyz=src_keypts[:, :, :3], points=None, returnidx=True)
    src_keyfeats = dv.index_points(src_feat, picked)
    src_cat = dv.Get_Cat_Feat_Src()(src_keypts, grouped, src_keyfeats)
    tgt_xyz = tgt_pts[:, :3, :].permute(0, 2, 1)
    _, tgt_feat = model.FE1(tgt_pts)
    R_rep = R_init.repeat(B, 1, 1)
    kT = src_keypts.permute(0, 2, 1)[:, :3, :]
    centres = torch.matmul(R_rep, kT.double()).permute(0, 2, 1)
    cand = dv.voxelize(centres, 1.0, 0.4)
    tgt_cat = dv.Get_Cat_Feat_Tgt()(cand, src_keypts, tgt_xyz, tgt_feat)
    s_dfe = model.DFE(src_cat, src=True).unsqueeze(2)
    t_dfe = model.DFE(tgt_cat, src=False).permute(0, 1, 3, 2)
    vcp = model.cpg(s_dfe, t_dfe, cand, 1.0, 0.4)
    return src_keypts[:, :, :3], vcp
