"""Sub-frame decomposition: how targets larger than 2048 px are rendered, and the multi-GPU shard unit.

The reference's guard band ends at device coordinate 2048 (`(2048 - half) / half`,
src/rgl/rglv/rglv_view_frustum.hxx:36-39, used as the whole clip extent :60-73), so neither it nor
this renderer accepts a target wider or taller than 2048 px.  A W x H frame is therefore rendered as
an nx x ny grid of sub-frames (4K = 2x2 x 1920x1080, 8K = 4x4), each with the projection
`crop(gx, gy) @ P`.  Sub-frames are independent: they are dealt round-robin to the ranks of one
node and only the resolved 8-bit images travel (gather to the presenting GPU).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

MAX_DIM = 2048


def crop_matrix(gx: int, gy: int, nx: int, ny: int) -> np.ndarray:
    """clip-space matrix that maps sub-window (gx, gy) of an nx x ny grid (gy = 0: top row) to the
    full NDC square: x' = nx x + (nx - 1 - 2 gx) w,  y' = ny y + (1 - ny + 2 gy) w"""
    m = np.eye(4, dtype=np.float32)
    m[0, 0] = nx
    m[0, 3] = nx - 1 - 2 * gx
    m[1, 1] = ny
    m[1, 3] = 1 - ny + 2 * gy
    return m


@dataclass(frozen=True)
class Subframe:
    index: int
    gx: int
    gy: int
    x0: int
    y0: int
    width: int
    height: int
    owner: int


class SubframePlan:
    """grid of equal sub-frames covering width x height, each <= max_w x max_h, owners round-robin"""

    def __init__(self, width: int, height: int, world_size: int = 1, max_w: int = 1920, max_h: int = 1080, owners=None):
        assert max_w <= MAX_DIM and max_h <= MAX_DIM
        self.width, self.height = width, height
        self.nx = -(-width // max_w)
        self.ny = -(-height // max_h)
        assert width % self.nx == 0 and height % self.ny == 0, "target must divide evenly into sub-frames"
        self.sub_w, self.sub_h = width // self.nx, height // self.ny
        assert self.sub_w % 4 == 0 and self.sub_h % 2 == 0
        self.world_size = world_size
        self.subframes = []
        for gy in range(self.ny):
            for gx in range(self.nx):
                i = gy * self.nx + gx
                owner = i % world_size if owners is None else int(owners[i])
                assert 0 <= owner < world_size
                self.subframes.append(Subframe(i, gx, gy, gx * self.sub_w, gy * self.sub_h, self.sub_w, self.sub_h, owner))

    @staticmethod
    def balance(costs, world_size: int):
        """owners for sub-frames of the given costs (e.g. measured device time): longest processing time first,
        each sub-frame to the currently least loaded rank -- sub-frames of a scene differ a lot in cost"""
        load = [0.0] * world_size
        owners = [0] * len(costs)
        for i in sorted(range(len(costs)), key=lambda k: -costs[k]):
            r = min(range(world_size), key=lambda k: (load[k], k))
            owners[i] = r
            load[r] += costs[i]
        return owners

    def owned_by(self, rank: int):
        return [s for s in self.subframes if s.owner == rank]

    def projection(self, proj: np.ndarray, s: Subframe) -> np.ndarray:
        return (crop_matrix(s.gx, s.gy, self.nx, self.ny) @ np.asarray(proj, np.float32)).astype(np.float32)

    def assemble(self, images: dict) -> np.ndarray:
        """images: {subframe index: (sub_h, sub_w) uint32} -> (height, width) uint32"""
        out = np.zeros((self.height, self.width), np.uint32)
        for s in self.subframes:
            out[s.y0:s.y0 + s.height, s.x0:s.x0 + s.width] = images[s.index]
        return out
