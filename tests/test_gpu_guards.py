"""Out-of-bounds writes, checked without compute-sanitizer (closed on this GPU pool, profiles/r02_sanitizer_unavailable.log):
every device buffer the forward allocates is placed between two guard regions filled with a byte pattern; after the
forward + pose solve (smoke size, a ModelNet-shaped batch, one KITTI-shaped full-size pair through the cluster FPS,
the indexed KNN and the tcgen05 embedding / CPG kernels) every guard must be intact, and the results must equal an
unguarded run (a kernel reading outside its inputs would see the pattern instead of the allocator's neighbours)."""
import importlib

import pytest
import torch

from conftest import PKG

pytestmark = pytest.mark.gpu
DEV = "cuda"
GUARD = 4096   # bytes on either side (multiple of 256: the payload keeps its alignment)


class GuardedAllocations:
    def __init__(self):
        self.bases = []
        self.orig = {}

    def _alloc(self, shape, dtype):
        n = 1
        for s in shape:
            n *= int(s)
        item = torch.empty((), dtype=dtype).element_size()
        pad = (-n * item) % 256
        base = torch.full((GUARD + n * item + pad + GUARD,), 0xA5, dtype=torch.uint8, device=DEV)
        self.bases.append((base, n * item))
        return base[GUARD:GUARD + n * item].view(dtype).view(*shape)

    def __enter__(self):
        for name in ("empty", "zeros", "empty_like", "zeros_like"):
            self.orig[name] = getattr(torch, name)

        def is_cuda(kw):
            d = kw.get("device", None)
            return d is not None and torch.device(d).type == "cuda"

        def empty(*shape, **kw):
            if not is_cuda(kw) or kw.get("pin_memory"):
                return self.orig["empty"](*shape, **kw)
            if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)):
                shape = tuple(shape[0])
            return self._alloc(shape, kw.get("dtype", torch.float32))

        def zeros(*shape, **kw):
            if not is_cuda(kw):
                return self.orig["zeros"](*shape, **kw)
            return empty(*shape, **kw).zero_()

        def empty_like(t, **kw):
            if not t.is_cuda or kw:
                return self.orig["empty_like"](t, **kw)
            return self._alloc(tuple(t.shape), t.dtype)

        def zeros_like(t, **kw):
            if not t.is_cuda or kw:
                return self.orig["zeros_like"](t, **kw)
            return self._alloc(tuple(t.shape), t.dtype).zero_()

        torch.empty, torch.zeros, torch.empty_like, torch.zeros_like = empty, zeros, empty_like, zeros_like
        return self

    def __exit__(self, *a):
        for name, fn in self.orig.items():
            setattr(torch, name, fn)

    def check(self):
        torch.cuda.synchronize()
        bad = 0
        for base, nbytes in self.bases:
            lo, hi = base[:GUARD], base[GUARD + nbytes + ((-nbytes) % 256):]
            bad += int((lo != 0xA5).sum()) + int((hi != 0xA5).sum())
        return len(self.bases), bad


@pytest.mark.parametrize("kind,B,N,G", [("modelnet", 1, 512, 5), ("modelnet", 3, 1024, 5), ("kitti", 1, 16384, 11),
                                         ("kitti", 2, 2048, 7)])
def test_forward_and_pose_write_only_inside_their_buffers(kind, B, N, G):
    dv = importlib.import_module(PKG)
    r = dv.synthetic.grid_radius(G)
    src, tgt, R, t = dv.synthetic.make_batch(kind, list(range(B)), N)
    torch.manual_seed(3)
    model = dv.DeepVCP(use_normal=src.shape[1] == 6, npoint=N, r=r, s=0.4).to(DEV).eval()
    g = torch.Generator().manual_seed(4)
    starts = (torch.randint(0, N, (B,), generator=g), torch.randint(0, 64, (B,), generator=g),
              torch.randint(0, N, (B,), generator=g))
    args = (src.to(DEV), tgt.to(DEV), R.to(DEV), torch.zeros(1, 3))
    kp0, vcp0 = model(*args, starts=starts, keep_stages=True)
    R0, t0 = dv.pose_from_forward(kp0, vcp0, R.to(DEV), t.view(B, 3, 1).to(DEV))
    torch.cuda.synchronize()
    with GuardedAllocations() as ga:
        kp1, vcp1 = model(*args, starts=starts, keep_stages=True)
        R1, t1 = dv.pose_from_forward(kp1, vcp1, R.to(DEV), t.view(B, 3, 1).to(DEV))
        nbuf, bad = ga.check()
    assert nbuf >= 15, "the guard did not see the forward's allocations (%d)" % nbuf
    assert bad == 0, "%d guard bytes overwritten" % bad
    assert torch.equal(kp0, kp1) and torch.equal(vcp0, vcp1) and torch.equal(R0, R1) and torch.equal(t0, t1)
