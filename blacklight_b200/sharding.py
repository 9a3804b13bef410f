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
