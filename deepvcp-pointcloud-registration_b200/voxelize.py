"""voxelize -- voxelize.py:19-83 of the reference: the G^3 candidate lattice
around every transformed key-point (origin (c - r) - s/2, step s, z fastest, no
sphere rejection), float64 arithmetic rounded to float32."""
from . import functional as F_


def voxelize(point_clouds, r, s):
    """point_clouds [B,N,3] -> [B,N,C,3] float32."""
    return F_.candidates(point_clouds, r, s)
