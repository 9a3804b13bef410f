"""Pixel sharding across GPUs: image rows are dealt round-robin over ranks because ray cost varies strongly
across the image (rays near the photon ring take several times more samples than edge rays); the grid is
replicated and the only exchange is the final gather of each rank's image rows."""
import numpy as np


def shard_rows(resolution, rank, world):
    """Rows owned by `rank` and the flat pixel indices m = row * res + col of its rays (reference pixel order,
    camera.cpp:393-396)."""
    rows = np.arange(rank, resolution, world)
    idx = (rows[:, None] * resolution + np.arange(resolution)[None, :]).ravel()
    return rows, idx


def assemble(parts, resolution, world):
    """Inverse of shard_rows for gathered per-rank images: parts[r] has shape (Q, rows_r * res)."""
    q = parts[0].shape[0]
    full = np.empty((q, resolution * resolution), dtype=parts[0].dtype)
    for r in range(world):
        _, idx = shard_rows(resolution, r, world)
        full[:, idx] = parts[r]
    return full


def shard_blocks(num_blocks, rank, world):
    """Adaptive refinement blocks owned by `rank` (contiguous block ids, round-robin)."""
    return np.arange(rank, num_blocks, world)


def gather_rows(mine, parts, full, resolution, rank, world, dist):
    """Final exchange of a row-sharded frame: every rank's (Q, rows_r * res) image part to rank 0 (one collective), where
    the rows are interleaved into `full` (Q, res, res).  Torch tensors on the device of the process group's backend
    (CUDA with nccl, CPU with gloo); parts / full are rank 0's receive buffers (None elsewhere).  A frame without image
    quantities (rendering only: Q = 0) has nothing to exchange.  Written so that no rank can leave it early: the only
    rank-dependent work is plain strided copies on rank 0."""
    if resolution % world != 0:   # the same on every rank: a gather needs equal parts
        raise ValueError('gather_rows: %d rows do not divide over %d ranks' % (resolution, world))
    if mine.shape[0] == 0:
        return full
    dist.gather(mine, parts if rank == 0 else None, dst=0)
    if rank == 0:
        q = mine.shape[0]
        for r in range(world):
            rows_r = len(range(r, resolution, world))
            full[:, r::world, :] = parts[r].reshape(q, rows_r, resolution)
    return full
