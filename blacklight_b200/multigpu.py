"""Adaptive ray tracing with the refinement blocks of every level sharded over GPUs.

Rays never interact and a block's refinement decision uses only its own pixels (reference
radiation_adaptive.cpp:76-85,120-125), so each rank traces and radiates a round-robin share of the blocks of a
level -- the ROOT level included, which is handed to the library block by block (bl_params.level0_block_major)
instead of as the full raster.  The only exchanges are, per level, an all-gather of the refinement flags
(bytes, so that every rank derives the identical child list: parents in index order x 4 children,
camera.cpp:445-459) and, at the end, a gather of each rank's image blocks.  The grid is replicated.

The algorithm is written once as a generator that yields its collectives; it is driven either by
torch.distributed (one process per GPU) or, for tests and single-GPU checks, by an in-process scheduler that
steps several virtual ranks in lock step.  Results are bitwise independent of the number of ranks.
"""
import numpy as np

from .sharding import shard_blocks


def _root_blocks(cfg):
    """Level-0 blocks in the reference's index order (block = v * nb + u) and, for each, the raster pixel indices
    of its bs x bs rays in block-major order."""
    res, bs = cfg.resolution, cfg.block_size
    nb = res // bs
    v, u = np.divmod(np.arange(nb * nb), nb)
    locs = np.stack([v, u], axis=1).astype(np.int32)
    i, j = np.divmod(np.arange(bs * bs), bs)
    pix = (v[:, None] * bs + i[None, :]) * res + u[:, None] * bs + j[None, :]   # (blocks, bs*bs)
    return locs, pix


def adaptive_worker(cfg, ctx, rank, world, max_level, num_render=0):
    """Generator.  Yields ('allgather', uint8 array) -> list of per-rank arrays, and finally
    ('gather', payload) -> list on rank 0 / None elsewhere; returns (via StopIteration.value) on rank 0 a list
    over levels of dict(locs=(B,2) int32, flags=(B,) uint8 or None, image=(Q, B*bs*bs) f64 in the reference's
    pixel order for that level [level 0: raster], stats=...), None on other ranks."""
    bs2 = cfg.block_size ** 2
    locs_all, pix = _root_blocks(cfg)
    pos_r, dir_r, fac_r = cfg.camera_root()
    mine_levels = []
    level = 0
    pos = dirs = fac = None
    while True:
        blocks = shard_blocks(len(locs_all), rank, world)
        if level == 0:
            sel = pix[blocks].ravel()
            pos, dirs, fac = pos_r[sel], dir_r[sel], fac_r[sel]
        else:
            sel = (blocks[:, None] * bs2 + np.arange(bs2)[None, :]).ravel()
            pos, dirs, fac = pos[sel], dirs[sel], fac[sel]
        stats = ctx.trace_level(level, pos, dirs, fac)
        image, render, rstats = ctx.radiate_level(level, num_render=num_render)
        flags_all = None
        if level < max_level:
            flags_mine, _ = ctx.refine_level(level, locs_all[blocks]) if len(blocks) else (np.zeros(0, np.uint8), 0)
            parts = yield ('allgather', flags_mine)
            flags_all = np.zeros(len(locs_all), np.uint8)
            for r, part in enumerate(parts):
                flags_all[shard_blocks(len(locs_all), r, world)] = part
        mine_levels.append(dict(blocks=blocks, image=image, render=render, locs=locs_all, flags=flags_all,
                                samples=rstats['num_samples'], bad=stats['num_bad_geodesics']))
        if flags_all is None or not flags_all.any():
            break
        level += 1
        # every rank derives the same child list and camera arrays; it then keeps its own share
        locs_all, pos, dirs, fac = cfg.camera_refined(level, locs_all, flags_all)
    gathered = yield ('gather', [dict(blocks=L['blocks'], image=L['image'], render=L['render']) for L in mine_levels])
    if gathered is None:
        return None
    out = []
    res = cfg.resolution
    for lv, L in enumerate(mine_levels):
        n_blocks = len(L['locs'])
        Q = L['image'].shape[0]
        full = np.empty((Q, n_blocks, bs2))
        for part in gathered:
            P = part[lv]
            full[:, P['blocks']] = P['image'].reshape(Q, len(P['blocks']), bs2)
        if lv == 0:   # back to the reference's raster order for the root level
            raster = np.empty((Q, res * res))
            raster[:, pix.ravel()] = full.reshape(Q, -1)
            image = raster
        else:
            image = full.reshape(Q, -1)
        out.append(dict(locs=L['locs'], flags=L['flags'], image=image))
    return out


def run_local(workers):
    """Drive several virtual ranks (generators from adaptive_worker) in lock step inside one process."""
    pending = [w.send(None) for w in workers]
    results = [None] * len(workers)
    live = list(range(len(workers)))
    while live:
        kinds = {pending[r][0] for r in live}
        assert len(kinds) == 1, 'ranks diverged: %s' % kinds
        kind = kinds.pop()
        payloads = [pending[r][1] for r in live]
        nxt = []
        for r in live:
            reply = payloads if kind == 'allgather' or r == 0 else None
            try:
                pending[r] = workers[r].send(reply)
                nxt.append(r)
            except StopIteration as stop:
                results[r] = stop.value
        live = nxt
    return results


def run_distributed(worker, rank, world):
    """Drive one rank's generator with torch.distributed object collectives (NCCL or gloo group)."""
    import torch.distributed as dist
    request = worker.send(None)
    while True:
        kind, payload = request
        if kind == 'allgather':
            parts = [None] * world
            dist.all_gather_object(parts, payload)
            reply = parts
        else:
            parts = [None] * world if rank == 0 else None
            dist.gather_object(payload, parts, dst=0)
            reply = parts
        try:
            request = worker.send(reply)
        except StopIteration as stop:
            return stop.value
