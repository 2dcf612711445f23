"""Pair sharding across GPUs: one process per GPU, pairs split into
contiguous blocks, no data-path collective; the only exchange is one all-gather
of the [B_local, 12] poses (R row-major 9 + t 3, float64) plus the timing
scalars bench.py reduces. The reference's only multi-device construct is
nn.DataParallel at batch size 1 (train.py:75-78), which cannot split work."""
import torch
import torch.distributed as dist


def shard_range(n_pairs: int, rank: int, world: int):
    """Contiguous block [lo, hi) of pair ids owned by `rank`; remainders go to the
    lowest ranks so block sizes differ by at most one."""
    base, rem = divmod(n_pairs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_poses(R: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """R [B,3,3], t [B,3,1] or [B,3] -> [B,12] float64."""
    B = R.shape[0]
    return torch.cat([R.reshape(B, 9).double(), t.reshape(B, 3).double()], dim=1).contiguous()


def unpack_poses(p: torch.Tensor):
    return p[:, :9].reshape(-1, 3, 3), p[:, 9:].reshape(-1, 3, 1)


def all_gather_poses(local: torch.Tensor, n_pairs: int, group=None) -> torch.Tensor:
    """local [B_local,12] -> [n_pairs,12] on every rank, in pair-id order. Ragged
    blocks are padded to the largest block for the collective and trimmed after."""
    if not dist.is_available() or not dist.is_initialized():
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(n_pairs, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros(width, local.shape[1], dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)


def register_shard(model, src, tgt, R_init, R_true, t_true, starts=None):
    """Forward + two-stage pose solve of this rank's pairs -> [B_local,12]."""
    from .deepVCP_loss import pose_from_forward
    kp, vcp = model(src, tgt, R_init, torch.zeros(1, 3), starts=starts)
    dev = kp.device
    R2, t2 = pose_from_forward(kp, vcp, R_true.to(dev), t_true.to(dev))
    return pack_poses(R2, t2)
