#!/usr/bin/env python
"""Auxiliary measurement for BASELINE.json configs[2]: example_adaptive parameters (polarized, tau image,
relative-Laplacian refinement) on a scaled root image, with the refinement blocks of every level sharded over the
ranks (blacklight_b200/multigpu.py).  Launch like bench.py:

  python tools/bench_adaptive.py --root 512 --levels 3 [--steps K]
  python -m torch.distributed.run --nproc-per-node N ... tools/bench_adaptive.py --root 512 --levels 3

Prints one JSON line on rank 0: rays of all levels per second (host wall clock between barriers; the trace and
radiate calls are synchronous), blocks and rays per level.  Not the contract bench (that is bench.py)."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--root', type=int, default=512)
    ap.add_argument('--levels', type=int, default=3)
    ap.add_argument('--block', type=int, default=8)
    ap.add_argument('--steps', type=int, default=2)
    ap.add_argument('--warmup', type=int, default=1)
    ap.add_argument('--region', type=float, default=6.0, help='half-width of a central window forced to the deepest level (0 = none)')
    args = ap.parse_args()
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    import torch
    import torch.distributed as dist
    import blacklight_b200 as bl
    from blacklight_b200 import multigpu
    from harness import Case
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    workdir = tempfile.mkdtemp(prefix='bl_adapt_%d_' % rank)
    over = {'camera_resolution': args.root, 'adaptive_max_level': args.levels, 'adaptive_block_size': args.block}
    if args.region > 0.0:
        w = '%g' % args.region
        over.update({'adaptive_num_regions': 1, 'adaptive_region_1_level': args.levels, 'adaptive_region_1_x_min': '-' + w,
                     'adaptive_region_1_x_max': w, 'adaptive_region_1_y_min': '-' + w, 'adaptive_region_1_y_max': w})
    case = Case(workdir, 'adaptive.input', over, threads=max(1, (os.cpu_count() or 1) // world))
    cfg = case.config(device=local_rank)
    cfg.set_level0_block_major(True)
    ctx = bl.Context(cfg)
    ctx.upload_grid(case.grid_arrays())

    def step():
        worker = multigpu.adaptive_worker(cfg, ctx, rank, world, args.levels)
        if world > 1:
            return multigpu.run_distributed(worker, rank, world)
        return multigpu.run_local([worker])[0]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        levels = step()
    barrier()
    dt = (time.perf_counter() - t0) / args.steps
    if rank == 0:
        bs2 = args.block ** 2
        blocks = [len(L['locs']) for L in levels]
        rays = [b * bs2 for b in blocks]
        flux = float(np.nanmean(levels[0]['image'][0]))
        print(json.dumps({'workload': 'example_adaptive parameters, root %d^2, block %d, up to %d levels, central window |x|,|y| < %g forced to the deepest level' % (args.root, args.block, args.levels, args.region),
                          'n_gpus': world, 'steps': args.steps, 'ms_per_step': 1e3 * dt, 'rays_per_s': sum(rays) / dt,
                          'blocks_per_level': blocks, 'rays_per_level': rays, 'mean_I_root': flux,
                          'rank0_stage_ms_last_step': {k: round(1e3 * v, 2) for k, v in multigpu.last_stage_seconds.items()},
                          'timing': 'host wall clock between barriers, includes host camera generation for refined levels, '
                                    'flag all-gather and final image gather'}))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
