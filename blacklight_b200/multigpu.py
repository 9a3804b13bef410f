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
import time

import numpy as np

from .sharding import shard_blocks

# host wall-clock seconds per stage of the last adaptive_worker run on this process (diagnostics for
# tools/bench_adaptive.py; trace / radiate / refine are synchronous calls, so they include the kernels)
last_stage_seconds = {}


_pinned_cache = {}


def _pinned(key, shape):
    """Pinned float64 host tensor of the given shape, kept between runs (pinning 200 MB costs more than copying it).
    The arrays a run returns are views of these buffers: valid until the next run on this process."""
    import torch
    t = _pinned_cache.get(key)
    if t is None or tuple(t.shape) != tuple(shape):
        t = torch.empty(shape, dtype=torch.float64).pin_memory()
        _pinned_cache[key] = t
    return t


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


def adaptive_worker(cfg, ctx, rank, world, max_level, num_render=0, device_images=None):
    """Generator.  Yields ('allgather', uint8 array) -> list of per-rank arrays, and finally
    ('gather_arrays', arrays, shapes of every rank's arrays) -> per-rank lists of arrays on rank 0 / None elsewhere; returns (via StopIteration.value) on rank 0 a list
    over levels of dict(locs=(B,2) int32, flags=(B,) uint8 or None, image=(Q, B*bs*bs) f64 in the reference's
    pixel order for that level [level 0: raster], stats=...), None on other ranks.
    device_images: callable (ctx, level) -> torch CUDA tensor (Q, rays) viewing the level's image in HBM (bl_device_image).
    With it nothing but the refinement flags touches the host before the end: every rank's blocks go to rank 0 device to
    device, are interleaved into the levels' images there, and each level is downloaded once."""
    if not getattr(ctx, 'level0_block_major', False):
        raise ValueError('adaptive_worker: the context must be created from a config with set_level0_block_major(True) '
                         '(the root level is handed over block by block)')
    if num_render > 0:
        raise ValueError('adaptive_worker: rendering is not gathered in sharded adaptive runs (use the C++ multi-device driver)')
    bs2 = cfg.block_size ** 2
    T = {k: 0.0 for k in ('camera', 'select', 'trace', 'radiate', 'refine', 'exchange', 'assemble')}
    last_stage_seconds.clear()
    last_stage_seconds.update(T)
    clock = time.perf_counter
    t0 = clock()
    locs_all, pix = _root_blocks(cfg)
    T['camera'] += clock() - t0
    mine_levels = []
    level = 0
    while True:
        t0 = clock()
        blocks = shard_blocks(len(locs_all), rank, world)
        t1 = clock()
        T['select'] += t1 - t0
        # the pixels of this rank's blocks are generated on the device (bl_trace_level_pixels): only the block list
        # crosses PCIe, and the host camera stage of round 1 is gone
        stats = ctx.trace_level_pixels(level, blocks=locs_all[blocks])
        t2 = clock()
        if device_images is None:
            image, render, rstats = ctx.radiate_level(level, num_render=num_render)
        else:
            _, render, rstats = ctx.radiate_level(level, num_render=num_render, download=False)
            image = device_images(ctx, level)
        t3 = clock()
        T['trace'] += t2 - t1
        T['radiate'] += t3 - t2
        flags_all = None
        if level < max_level:
            flags_mine, _ = ctx.refine_level(level, locs_all[blocks]) if len(blocks) else (np.zeros(0, np.uint8), 0)
            t4 = clock()
            parts = yield ('allgather', flags_mine)
            t5 = clock()
            flags_all = np.zeros(len(locs_all), np.uint8)
            for r, part in enumerate(parts):
                flags_all[shard_blocks(len(locs_all), r, world)] = part
            T['refine'] += t4 - t3
            T['exchange'] += t5 - t4
        mine_levels.append(dict(blocks=blocks, image=image, render=render, locs=locs_all, flags=flags_all,
                                samples=rstats['num_samples'], bad=stats['num_bad_geodesics']))
        if flags_all is None or not flags_all.any():
            break
        level += 1
        # every rank derives the same child list (parents in index order x 4 children (2v..2v+1) x (2u..2u+1),
        # camera.cpp:445-459) and then builds the camera arrays of its own share
        t0 = clock()
        parents = locs_all[flags_all != 0]
        dv, du = np.array([0, 0, 1, 1], np.int32), np.array([0, 1, 0, 1], np.int32)
        locs_all = np.stack([2 * parents[:, None, 0] + dv[None, :], 2 * parents[:, None, 1] + du[None, :]], axis=2)
        locs_all = np.ascontiguousarray(locs_all.reshape(-1, 2), np.int32)
        T['camera'] += clock() - t0
    t0 = clock()
    # final exchange: every rank's image blocks of every level to rank 0, as flat float64 arrays whose shapes all
    # ranks can derive (levels x Q x this rank's blocks x bs2) -- no pickling, NCCL point-to-point when distributed
    Q = mine_levels[0]['image'].shape[0]
    shapes = [[(Q, len(shard_blocks(len(L['locs']), r, world)) * bs2) for L in mine_levels] for r in range(world)]
    on_device = device_images is not None
    gathered = yield ('gather_arrays', [L['image'] if on_device else np.ascontiguousarray(L['image'], np.float64) for L in mine_levels],
                      shapes)
    T['exchange'] += clock() - t0
    last_stage_seconds.update(T)
    if gathered is None:
        return None
    t0 = clock()
    out = []
    res = cfg.resolution
    if on_device:
        import torch
        dev = mine_levels[0]['image'].device
        pix_index = torch.from_numpy(np.ascontiguousarray(pix.ravel(), np.int64)).to(dev)
    for lv, L in enumerate(mine_levels):
        n_blocks = len(L['locs'])
        full = torch.empty((Q, n_blocks, bs2), dtype=torch.float64, device=dev) if on_device else np.empty((Q, n_blocks, bs2))
        for r, part in enumerate(gathered):   # shard_blocks deals blocks round-robin: rank r owns blocks r, r + world, ...
            full[:, r::world] = part[lv].reshape(Q, -1, bs2)
        if lv == 0:   # back to the reference's raster order for the root level
            raster = torch.empty((Q, res * res), dtype=torch.float64, device=dev) if on_device else np.empty((Q, res * res))
            raster[:, pix_index if on_device else pix.ravel()] = full.reshape(Q, -1)
            image = raster
        else:
            image = full.reshape(Q, -1)
        if on_device:
            host = _pinned(('level', lv), tuple(image.shape))   # one device-to-host copy per level, into pinned memory
            host.copy_(image)
            image = host.numpy()
        out.append(dict(locs=L['locs'], flags=L['flags'], image=image))
    T['assemble'] += clock() - t0
    last_stage_seconds.update(T)
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
    import torch
    import torch.distributed as dist
    request = worker.send(None)
    while True:
        kind, payload = request[0], request[1]
        if kind == 'allgather':
            parts = [None] * world
            dist.all_gather_object(parts, payload)
            reply = parts
        elif kind == 'gather_arrays':
            # one flat float64 message per rank; device tensors with the nccl backend (NVLink), host tensors with gloo
            shapes = request[2]
            dev = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend() == 'nccl' else torch.device('cpu')
            count = lambda r: sum(int(np.prod(sh)) for sh in shapes[r])
            tensors = len(payload) > 0 and torch.is_tensor(payload[0])   # device-resident images: nothing goes through the host
            if rank == 0:
                starts = np.concatenate([[0], np.cumsum([count(r) for r in range(1, world)])]).astype(np.int64)
                buf = torch.empty(int(starts[-1]), dtype=torch.float64, device=dev)
                ops = [dist.P2POp(dist.irecv, buf[int(starts[r - 1]):int(starts[r])], r) for r in range(1, world) if count(r)]
                if ops:
                    for q in dist.batch_isend_irecv(ops):
                        q.wait()
                flat_all = buf if tensors else buf.cpu().numpy()          # one device-to-host copy for all ranks
                reply = [payload]
                for r in range(1, world):
                    arrays, at = [], int(starts[r - 1])
                    for sh in shapes[r]:
                        n = int(np.prod(sh))
                        arrays.append(flat_all[at:at + n].reshape(sh))
                        at += n
                    reply.append(arrays)
            else:
                if count(rank):
                    if tensors:
                        flat = torch.cat([a.reshape(-1) for a in payload])
                    else:
                        flat = torch.from_numpy(np.concatenate([np.asarray(a, np.float64).ravel() for a in payload])).to(dev)
                    for q in dist.batch_isend_irecv([dist.P2POp(dist.isend, flat, 0)]):
                        q.wait()
                reply = None
        else:
            parts = [None] * world if rank == 0 else None
            dist.gather_object(payload, parts, dst=0)
            reply = parts
        try:
            request = worker.send(reply)
        except StopIteration as stop:
            return stop.value
