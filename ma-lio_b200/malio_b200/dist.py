"""Multi-GPU host logic: one process per GPU, the scan sharded by contiguous point block.

The map snapshot, the pose tables and the filter state are replicated; per measurement pass the ranks exchange
  (1) MIN over {min_u, -max_u, min_tau, -max_tau}            (laserMapping.cpp:615-628, 700-703)
  (2) SUM over the reduced system, the 3x3 normal scatter and N_eff
both inside libmalio_b200.so on NCCL (malio_comm_init).  torch.distributed only carries the 128-byte NCCL id.
"""
from __future__ import annotations

import numpy as np

from . import capi


def shard_bounds(n_points: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of the merged scan owned by `rank`; blocks differ by at most one point."""
    assert 0 <= rank < world
    return (n_points * rank) // world, (n_points * (rank + 1)) // world


def localization_weight(S6: np.ndarray, params: capi.Params) -> tuple[float, np.ndarray]:
    """laserMapping.cpp:745-756 from the (summed) 3x3 scatter of the weighted normals."""
    S = np.array([[S6[0], S6[1], S6[2]], [S6[1], S6[3], S6[4]], [S6[2], S6[4], S6[5]]])
    sv = np.sqrt(np.clip(np.sort(np.linalg.eigvalsh(S))[::-1], 0, None))
    w = sv[2] / sv[0]
    if w > params.localize_thresh_max:
        w = params.localize_cov_max
    elif w < params.localize_thresh_min:
        w = params.localize_cov_min
    else:
        w = (params.localize_cov_max - params.localize_cov_min) * (w - params.localize_thresh_min) / \
            (params.localize_thresh_max - params.localize_thresh_min) + params.localize_cov_min
    return float(w), sv


def init_comm(model, rank: int, world: int, device=None):
    """Create the NCCL communicator of `model` (a plugin.MeasurementModel): rank 0 makes the unique id, the id
    travels through the default torch.distributed group."""
    import torch
    import torch.distributed as dist
    uid = np.zeros(capi.NCCL_UNIQUE_ID_BYTES, np.uint8)
    if rank == 0:
        uid = type(model).nccl_unique_id()
    t = torch.from_numpy(uid.copy())
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, 0)
    model.comm_init(t.cpu().numpy(), rank, world)
