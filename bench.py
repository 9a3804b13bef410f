#!/usr/bin/env python
"""Headline benchmark: camera rays/sec of the per-pixel hot path (geodesics + sampling + coefficients + transfer).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                     times the unmodified reference's CPU path

Default workload = BASELINE.json configs[3], the configuration the metric is quoted on: the mock Athena++ snapshot
rendered as a 4096^2 polarized (Stokes IQUV) image, kappa-distribution electrons (kappa = 4), 4 frequencies, with the
image rows dealt round-robin over the N ranks (STRONG scaling: the frame is fixed, per-GPU work shrinks as N grows).
A step is one pass of the hot path over the whole frame.  `value` is rays/s with the camera arrays and the grid already
resident in HBM; `e2e` is the same metric through the C ABI from pinned HOST buffers: H2D of the step's input (the list of image rows
this rank renders -- the camera pixels are generated on the device; --host-camera uploads host-built arrays), the
kernels, the gather of every rank's rows into rank 0's image over NCCL, and the D2H of the assembled frame, all inside
the timed region.  The other configurations of BASELINE.json (formula plasma, 1024^2 unpolarized, adaptive refinement
sharded over the ranks, true colour, false-colour rendering) are measured the same way with fewer steps and appended
under `extra`.  One JSON line on rank 0.

Rooflines are in EXECUTED FP64 operations: profiles/executed_flops.json holds, per workload and kernel, the thread-level
DADD + DMUL + 2 DFMA count per unit of work measured with ncu (tools/ncu_capture.sh regenerates it and records the hash
of the kernel sources it was taken from); here that count is multiplied by the live unit count of the run and divided by
the kernel's live CUDA-event time and by the DFMA peak measured in the same run.
"""
import argparse
import datetime
import hashlib
import json
import math
import os
import shutil
import subprocess
import sys
import tempfile
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

REF_BIN = os.path.join(ROOT, 'oracle', '_ref', 'blacklight')   # the checker's build of the unmodified reference
EXECUTED_JSON = os.path.join(ROOT, 'profiles', 'executed_flops.json')
KERNEL_SOURCES = ['geodesic_dp.cu', 'ks_exact.cuh', 'glibc_math.cuh', 'radiate_unpol.cu', 'radiate_pol.cu', 'radiate_pol_split.cu',
                  'pol_common.cuh', 'rad_sample.cuh', 'bf_math.cuh']

# Algorithmic bytes per sample: one 64-byte step-buffer record written by
# the geodesic kernel and read once by the radiation kernels (the 256-byte trilinear gather is served by L2 for the
# 20 MB mock grid)
RECORD_BYTES_PER_SAMPLE = 64.0

C4 = {'image_polarization': 'true', 'image_num_frequencies': 4, 'image_frequency_start': '8.6e10',
      'image_frequency_end': '3.45e11', 'image_frequency_spacing': 'log', 'plasma_kappa_frac': '1.0',
      'plasma_kappa': '4.0', 'plasma_w': '1.0'}
WORKLOADS = {
    # name: (input template, overrides, default image side, description)
    'c4': ('simulation.input', C4, 4096,
           'BASELINE configs[3]: mock Athena++ snapshot, polarized Stokes IQUV, kappa = 4 electrons, 4 frequencies 86-345 GHz'),
    'polarized_thermal': ('simulation.input', {'image_polarization': 'true'}, 1024,
                          'mock Athena++ snapshot, polarized thermal synchrotron, 1 frequency'),
    'simulation': ('simulation.input', {}, 1024,
                   'BASELINE configs[1]: mock Athena++ snapshot, example_simulation parameters (unpolarized thermal synchrotron, trilinear)'),
    'formula': ('formula.input', {}, 512, 'BASELINE configs[0]: example_formula parameters (formula plasma, a = 0.9)'),
    'true_color': ('true_color.input', {}, 512, 'BASELINE configs[4]: example_true_color parameters (10 frequencies, unpolarized)'),
    'render': ('render.input', {}, 1024, 'BASELINE configs[4]: example_render parameters (false-colour rendering, flat space)'),
    'adaptive': ('adaptive.input', {}, 512,
                 'BASELINE configs[2]: example_adaptive parameters (polarized + tau, relative-Laplacian refinement), root image '
                 'scaled up, 3 levels, central window forced to the deepest level, refinement blocks sharded over the ranks'),
}
WORKLOADS['polarized'] = WORKLOADS['c4']
EXTRAS = ['simulation', 'formula', 'adaptive', 'true_color', 'render']


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='c4', choices=sorted(WORKLOADS))
    ap.add_argument('--resolution', type=int, default=0, help='image side (0 = the workload\'s default)')
    ap.add_argument('--scaling', default='strong', choices=['strong', 'weak'],
                    help='strong: the frame is fixed and its rows are dealt over the ranks; weak: the frame grows so that '
                         'every rank keeps resolution^2 rays')
    ap.add_argument('--tile-rays', type=int, default=0)
    ap.add_argument('--grid-scale', type=int, default=1,
                    help='refine the mock snapshot by this factor per dimension (4: 308x256x512 cells, 1.3 GB of primitives -- '
                         'the gather leaves L2 and becomes HBM traffic; SURVEY.md section 8d)')
    ap.add_argument('--cpu-resolution', type=int, default=0, help='side of the bounded CPU sample (0 = auto)')
    ap.add_argument('--e2e-budget-s', type=float, default=60.0,
                    help='the end-to-end leg runs over min(steps, budget / seconds per pass) passes, at least 3')
    ap.add_argument('--host-camera', action='store_true',
                    help='end-to-end leg: build the camera arrays on the host and upload them (round 1) instead of generating '
                         'the pixels on the device')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='only the main workload (no `extra` lines)')
    ap.add_argument('--dump-units', default='', help='write the unit counts of the main workload (for tools/ncu_flops_json.py)')
    return ap.parse_args()


def make_case(name, workdir, resolution, grid_scale=1, extra_over=None):
    from blacklight_b200.cases import Case
    base, over, _, _ = WORKLOADS[name]
    over = dict(over)
    over['camera_resolution'] = resolution
    over.update(extra_over or {})
    mock = None
    if grid_scale > 1 and base != 'formula.input':
        mock = {'n_r': 77 * grid_scale, 'n_th': 64 * grid_scale, 'n_ph': 128 * grid_scale}
    # host threads (camera pixels, reader conversions) per rank: the box's cores shared among the ranks
    world = int(os.environ.get('WORLD_SIZE', '1'))
    return Case(workdir, base, over, mock=mock, threads=max(1, (os.cpu_count() or 1) // world))


def source_hash():
    h = hashlib.sha256()
    for f in KERNEL_SOURCES:
        with open(os.path.join(ROOT, 'blacklight_b200', 'csrc', f), 'rb') as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def load_executed():
    try:
        d = json.load(open(EXECUTED_JSON))
    except (OSError, ValueError):
        return {'entries': {}, 'missing': True}
    d['stale'] = d.get('source_hash') != source_hash()
    return d


class ClockSampler:
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.path = tempfile.mktemp(suffix='.csv')
        self.proc = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.QUERY,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(',')]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(names, f[5:9]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            out['sm_mhz'] = float(np.median(sm))
            out['sm_max_mhz'] = float(max(mx))
        out['reasons'] = sorted(reasons)
        return out


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the unmodified reference (oracle/_ref/blacklight) on the host cores

def reference_rays_per_s(case, threads):
    """Run the unmodified reference once on `case`; rays/s from its own timers (geodesic + sample + image + render)."""
    from blacklight_b200.cases import parse_timers, write_input
    kv = dict(case.kv)
    kv['num_threads'] = str(threads)
    out = os.path.join(case.dir, 'out_cpu')
    os.makedirs(out, exist_ok=True)
    kv['output_file'] = os.path.join(out, 'image.npz')
    path = os.path.join(case.dir, 'cpu.input')
    write_input(path, kv)
    proc = subprocess.run([REF_BIN, path], cwd=case.dir, capture_output=True, text=True, timeout=3600)
    if proc.returncode != 0 or 'Calculation completed' not in proc.stdout:
        raise RuntimeError('reference failed: ' + proc.stdout[-2000:] + proc.stderr[-2000:])
    t = parse_timers(proc.stdout)
    compute = t.get('Integrating geodesics', 0.0) + t.get('Sampling simulation', 0.0) + t.get('Integrating image', 0.0) \
        + t.get('Rendering', 0.0)
    rays = int(case.kv['camera_resolution']) ** 2
    return rays / compute, compute


def auto_cpu_resolution(name, threads):
    # rays/s of the reference per thread measured on this pool (16 threads: ~11 k unpolarized, ~0.7 k C4); ~15 s of work
    per_thread = {'simulation': 700.0, 'formula': 350.0, 'c4': 45.0, 'polarized': 45.0, 'polarized_thermal': 250.0,
                  'true_color': 250.0, 'render': 1500.0, 'adaptive': 250.0}[name]
    side = int(math.sqrt(per_thread * threads * 15.0))
    return max(24, min(256, side // 8 * 8))


def workload_label(name, res_total, n_gpus, scaling):
    return '%s; image %dx%d over %d GPU(s), %s scaling' % (WORKLOADS[name][3], res_total, res_total, n_gpus, scaling)


def frame_side(args, name, n_gpus):
    res = args.resolution or WORKLOADS[name][2]
    if args.scaling == 'weak' and n_gpus > 1:
        res = int(math.ceil(res * math.sqrt(n_gpus) / (8 * n_gpus))) * 8 * n_gpus
    return res


ADAPTIVE_SAMPLE = {'image_polarization': 'true', 'image_tau': 'true'}   # the per-ray physics of example_adaptive


def run_reference_arm(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    name = args.workload
    res_total = frame_side(args, name, args.gpus)
    line = {'impl': 'reference', 'metric': 'camera rays/sec (geodesic + polarized RT)', 'unit': 'rays/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'higher_is_better': True, 'scaling': args.scaling,
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic'}
    if not os.path.exists(REF_BIN):
        emit({'impl': 'reference', 'unavailable': 'oracle/_ref/blacklight was not built (no /root/reference at build time)'})
        return
    side = args.cpu_resolution or auto_cpu_resolution(name, threads)
    workdir = tempfile.mkdtemp(prefix='bl_ref_')
    try:
        if name == 'adaptive':   # the reference's per-ray cost of that physics on a plain frame
            case = make_case('simulation', workdir, side, 1, ADAPTIVE_SAMPLE)
        else:
            case = make_case(name, workdir, side, 1)
        computes = []
        steps, warmup = args.steps, args.warmup
        i = 0
        while i < warmup + steps:
            _, c = reference_rays_per_s(case, threads)
            if i >= warmup:
                computes.append(c)
            if i == 0 and c * (warmup + steps) > 240.0:   # keep the whole run within a few minutes
                warmup = 0
                steps = max(1, min(steps, int(240.0 / c)))
                computes = [c]
            i += 1
        value = side * side * len(computes) / sum(computes)
        line.update({'value': value, 'ms_per_step': 1e3 * sum(computes) / len(computes), 'steps': len(computes), 'warmup': warmup,
                     'config': {'workload': workload_label(name, res_total, args.gpus, args.scaling),
                                'note': 'the reference CPU path timed on a bounded sample (%dx%d rays) of the same camera and '
                                        'physics; rays/s is resolution independent' % (side, side)},
                     'cpu_baseline': {'value': value, 'unit': 'rays/s', 'cores': threads, 'kind': 'reference',
                                      'sample': '%dx%d rays of the same camera; reference timers geodesic + sample + image'
                                                % (side, side)},
                     'e2e': {'value': value, 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}})
        emit(line)
    finally:
        shutil.rmtree(workdir, ignore_errors=True)


_JSON_FD = None


def emit(line):
    """Write the one JSON result line to the process's ORIGINAL stdout (see main)."""
    data = (json.dumps(line) + '\n').encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


# ---------------------------------------------------------------------------------------------------------------------
# one workload on this rank's GPU

class _DeviceArray:
    """__cuda_array_interface__ view of a device buffer owned by the library."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {'shape': tuple(shape), 'typestr': '<f8', 'data': (int(ptr), False), 'version': 2}


def kernel_rooflines(name, st, stages, ms_geo, ms_rad, fp64_peak, hbm_peak, executed, F, polarized):
    """Per-kernel live times and executed-FP64 rooflines of one step on rank 0."""
    alias = {'polarized': 'c4', 'adaptive': 'polarized_thermal'}   # same kernels, same per-sample work
    entry = executed.get('entries', {}).get(alias.get(name, name), {})
    units = {'attempt': st['num_attempts'], 'sample': st['num_samples'], 'ray': st['num_rays']}
    kernels = []

    def add(kname, ms, bytes_algo):
        if ms <= 0:
            return
        k = {'kernel': kname, 'ms': ms, 'bound': 'fp64', 'peak': fp64_peak, 'unit': 'TFLOP/s'}
        e = entry.get(kname)
        if e:
            flop = e['flop_per_unit'] * units[e['unit']]
            k.update({'achieved': flop / (ms * 1e-3) / 1e12, 'flop_per_unit': e['flop_per_unit'], 'per': e['unit'],
                      'fp64_pipe_active_ncu': e.get('fp64_pipe_active'),
                      'traffic': e['dram_bytes_per_unit'] * units[e['unit']] if e.get('dram_bytes_per_unit') else None})
            k['frac'] = k['achieved'] / fp64_peak if fp64_peak else None
        else:
            k.update({'achieved': None, 'frac': None, 'traffic': None,
                      'note': 'no ncu flop count for this kernel / workload in profiles/executed_flops.json'})
        k['hbm'] = {'achieved': bytes_algo / (ms * 1e-3) / 1e9, 'peak': hbm_peak, 'unit': 'GB/s', 'algorithmic_bytes': bytes_algo}
        kernels.append(k)

    ns = st['num_samples']
    add('geodesic_dp_kernel', ms_geo, ns * RECORD_BYTES_PER_SAMPLE)
    if stages and stages['slab'] > 0:
        # scratch between the stages: 8 + 7 doubles per sample out of the sampling stage, 8 read + 11 written by the
        # geometry stage, 7 read + 8 F written by the coefficient stage, 11 + 8 F read by the transfer stage
        # (radiate_pol_split.cu); the first two read the 64-byte step-buffer record
        add('pol_sampling_kernel', stages['sampling_ms'], ns * (RECORD_BYTES_PER_SAMPLE + 15 * 8.0))
        add('pol_geometry_kernel', stages['geometry_ms'], ns * (RECORD_BYTES_PER_SAMPLE + 19 * 8.0))
        add('pol_coefficient_kernel', stages['coefficients_ms'], ns * (7 + 8 * F) * 8.0)
        add('pol_transfer_kernel', stages['transfer_ms'], ns * (11 + 8 * F) * 8.0)
    else:
        add('radiate_polarized_kernel' if polarized else 'radiate_unpolarized_kernel', ms_rad, ns * RECORD_BYTES_PER_SAMPLE)
    return kernels


def measure(args, name, resolution, steps, warmup, rank, world, local_rank, main):
    """Time `steps` passes of workload `name` (rows of a resolution^2 frame dealt over the ranks).  Returns the raw
    result dict on rank 0 (None elsewhere)."""
    import torch
    import torch.distributed as dist
    import blacklight_b200 as bl
    from blacklight_b200.sharding import gather_rows, shard_rows
    dev = torch.device('cuda', local_rank)
    workdir = tempfile.mkdtemp(prefix='bl_bench_%d_' % rank)
    try:
        case = make_case(name, workdir, resolution, args.grid_scale)
        cfg = case.config(device=local_rank, tile_rays=args.tile_rays)
        F = int(cfg.keys.get('image_num_frequencies', '1'))
        R = int(cfg.keys.get('render_num_images', '0')) if case.sim else 0
        polarized = case.sim and cfg.keys.get('image_polarization', 'false') == 'true'
        # this rank's rays: image rows rank, rank + world, ... (cost varies strongly across the image).  Their camera
        # pixels are generated on the device from the row list (bl_trace_level_pixels); --host-camera builds the camera
        # arrays on the host instead and uploads them from pinned memory every step, as round 1 did (72 bytes per ray)
        rows, idx = shard_rows(resolution, rank, world)
        n_rays = len(idx)
        rows_np = torch.from_numpy(rows.astype(np.int32)).pin_memory().numpy()
        pos_np = dir_np = fac_np = None
        if args.host_camera:
            pos_all, dir_all, fac_all = cfg.camera_root()
            pos = torch.from_numpy(pos_all[idx]).pin_memory()
            dirs = torch.from_numpy(dir_all[idx]).pin_memory()
            fac = torch.from_numpy(fac_all[idx]).pin_memory()
            del pos_all, dir_all, fac_all
            pos_np, dir_np, fac_np = pos.numpy(), dirs.numpy(), fac.numpy()
        h2d_bytes = n_rays * 72 if args.host_camera else rows_np.nbytes
        del idx

        ctx = bl.Context(cfg)
        info = ctx.device_info()
        Q = ctx.num_quantities
        # gather buffers first, so that the library sizes its waves from what is really free
        image_host = torch.empty((Q, resolution * resolution if rank == 0 else n_rays), dtype=torch.float64).pin_memory()
        render_host = np.empty((R, 3, n_rays)) if R > 0 else None
        full_dev, parts = None, None
        if world > 1 and rank == 0:
            full_dev = torch.empty((Q, resolution, resolution), dtype=torch.float64, device=dev)
            parts = [torch.empty((Q, len(shard_rows(resolution, r, world)[1])), dtype=torch.float64, device=dev) for r in range(world)]
        if case.sim:
            ctx.upload_grid(case.grid_arrays())

        # every kernel and copy of the library is issued on the context's own stream: time with CUDA events recorded
        # on THAT stream (torch's current stream sees none of it)
        lib_stream = torch.cuda.ExternalStream(ctx.cuda_stream(), device=dev)

        def barrier():
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        def timed(fn, count):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            t0 = time.perf_counter()
            ev0.record(lib_stream)
            out = None
            for _ in range(count):
                out = fn()
            lib_stream.wait_stream(torch.cuda.current_stream())   # NCCL gather, assembly and D2H of the last step
            ev1.record(lib_stream)
            ev1.synchronize()
            barrier()
            return ev0.elapsed_time(ev1) * 1e-3, time.perf_counter() - t0, out

        def step_e2e():
            if args.host_camera:
                ctx.trace_level(0, pos_np, dir_np, fac_np)                 # H2D of the camera arrays (+ trace if resident)
            else:
                ctx.trace_level_pixels(0, rows=rows_np)                    # H2D of the row list, pixels on the device (+ trace)
            if world == 1:
                _, _, st_ = ctx.radiate_level(0, image=image_host.numpy(), render=render_host, num_render=R)   # kernels + D2H
                return st_
            _, _, st_ = ctx.radiate_level(0, download=False, render=render_host, num_render=R)
            if Q > 0:
                ptr, shape = ctx.device_image(0)
                mine = torch.as_tensor(_DeviceArray(ptr, shape), device=dev)
                gather_rows(mine, parts, full_dev, resolution, rank, world, dist)   # rows of every rank to rank 0 over NVLink
                if rank == 0:
                    image_host.copy_(full_dev.view(Q, -1), non_blocking=True)    # one D2H of the assembled frame
            torch.cuda.current_stream().synchronize()
            return st_

        ms_acc = {'geo': 0.0, 'rad': 0.0, 'stage': [0.0, 0.0, 0.0, 0.0], 'slab': 0}

        def step_resident():
            st0 = ctx.retrace_level(0)
            _, _, st_ = ctx.radiate_level(0, download=False)
            # a resident level is traced by retrace_level, a level in waves inside radiate_level
            ms_acc['geo'] += st0['ms_geodesic'] if st0['ms_geodesic'] > 0 and st_['ms_geodesic'] == st0['ms_geodesic'] else st_['ms_geodesic']
            ms_acc['rad'] += st_['ms_radiation']
            sg = ctx.polarized_stage_ms(0)
            for k, key in enumerate(('geometry_ms', 'coefficients_ms', 'transfer_ms', 'sampling_ms')):
                ms_acc['stage'][k] += sg[key]
            ms_acc['slab'] = sg['slab']
            return st_

        t_w = time.perf_counter()
        for _ in range(warmup):
            step_e2e()
        torch.cuda.synchronize()
        step_s = (time.perf_counter() - t_w) / max(1, warmup)
        # `value` is always timed over exactly `steps` passes.  The end-to-end leg repeats the same passes with the host
        # copies inside; when one pass takes seconds (the 4096^2 frame on one GPU) it runs over fewer passes (at least 3,
        # stated in e2e.steps) so that the whole default run still ends within minutes.
        e2e_steps = steps
        if step_s * steps > args.e2e_budget_s:
            e2e_steps = max(min(3, steps), min(steps, int(args.e2e_budget_s / step_s)))
        if world > 1:   # every rank must run the same number of passes (the gather is a collective)
            es = torch.tensor([e2e_steps], dtype=torch.int64, device=dev)
            dist.all_reduce(es, op=dist.ReduceOp.MIN)
            e2e_steps = int(es.item())
        fp64_peak = ctx.measure_fp64_peak() if main else None
        sampler = ClockSampler(local_rank) if main else None
        t_e2e, wall_e2e, _ = timed(step_e2e, e2e_steps)
        launches0 = ctx.launch_count()
        t_res, wall_res, st = timed(step_resident, steps)
        launches = ctx.launch_count() - launches0
        clocks = sampler.stop() if sampler else None

        times = torch.tensor([t_e2e, t_res], dtype=torch.float64, device=dev)
        counts = torch.tensor([float(n_rays), float(launches), float(st['num_samples']), float(h2d_bytes)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(times, op=dist.ReduceOp.MAX)
            dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        t_e2e, t_res = times.tolist()
        total_rays, total_launches, total_samples, total_h2d = counts.tolist()
        out = None
        if rank == 0:
            K = steps
            img = image_host.numpy()
            finite = np.isfinite(img)
            out = {
                'workload': name, 'value': total_rays * K / t_res, 'unit': 'rays/s', 'ms_per_step': 1e3 * t_res / K,
                'e2e': {'value': total_rays * e2e_steps / t_e2e, 'unit': 'rays/s', 'ms_per_step': 1e3 * t_e2e / e2e_steps,
                        'steps': e2e_steps, 'h2d_bytes_per_step': int(total_h2d), 'd2h_bytes_per_step': int(total_rays * 8 * Q),
                        'camera': 'host arrays uploaded (72 B per ray)' if args.host_camera else
                                  'pixels generated on the device from the row list (bl_trace_level_pixels)'},
                'gpu_launches': int(total_launches), 'rays': int(total_rays), 'rays_rank0': n_rays, 'frequencies': F,
                'samples_per_step': int(total_samples), 'ray_freq_per_s': total_rays * K / t_res * F,
                # the assembled frame of the last end-to-end step: equal across N means the sharded image is bitwise the same
                'image_crc32': '%08x' % (zlib.crc32(img.tobytes()) & 0xffffffff),
                'image_sum': float(img[finite].sum()) if img.size else 0.0,
                'host_wall_ms_per_step': 1e3 * wall_res / K, 'host_wall_ms_per_step_e2e': 1e3 * wall_e2e / e2e_steps,
                'l2': 'inputs larger than L2: %.1f GB step buffer written and re-read per step on rank 0' % (st['num_samples'] * 64 / 1e9),
                '_st': st, '_ms': {'geo': ms_acc['geo'] / K, 'rad': ms_acc['rad'] / K}, '_polarized': polarized,
                '_stages': {'geometry_ms': ms_acc['stage'][0] / K, 'coefficients_ms': ms_acc['stage'][1] / K,
                            'transfer_ms': ms_acc['stage'][2] / K, 'sampling_ms': ms_acc['stage'][3] / K, 'slab': ms_acc['slab']},
                '_fp64_peak': fp64_peak, '_clocks': clocks, '_device': info['name'], '_resolution': resolution,
                '_passes': warmup + e2e_steps + steps,
            }
        ctx.close()
        del full_dev, parts
        torch.cuda.empty_cache()
        return out
    finally:
        shutil.rmtree(workdir, ignore_errors=True)


def measure_adaptive(root, levels, steps, warmup, rank, world, local_rank):
    """BASELINE configs[2]: adaptive refinement with the blocks of every level sharded over the ranks
    (blacklight_b200/multigpu.py).  Host wall clock between barriers: the run interleaves host camera generation, the
    flag all-gather and the final gather with the kernels, so it is end to end by construction."""
    import torch
    import torch.distributed as dist
    import blacklight_b200 as bl
    from blacklight_b200 import multigpu
    workdir = tempfile.mkdtemp(prefix='bl_adapt_%d_' % rank)
    try:
        w = '6'
        over = {'adaptive_max_level': levels, 'adaptive_block_size': 8, 'adaptive_num_regions': 1, 'adaptive_region_1_level': levels,
                'adaptive_region_1_x_min': '-' + w, 'adaptive_region_1_x_max': w, 'adaptive_region_1_y_min': '-' + w,
                'adaptive_region_1_y_max': w}
        case = make_case('adaptive', workdir, root, 1, over)
        cfg = case.config(device=local_rank)
        cfg.set_level0_block_major(True)
        ctx = bl.Context(cfg)
        ctx.upload_grid(case.grid_arrays())

        dev = torch.device('cuda', local_rank)
        Q = ctx.num_quantities

        def device_image(c, level):   # the level's image where bl_radiate_level left it in HBM
            if c._rays.get(level, 0) == 0:   # this rank owns no block of the level
                return torch.empty((Q, 0), dtype=torch.float64, device=dev)
            ptr, shape = c.device_image(level)
            return torch.as_tensor(_DeviceArray(ptr, shape), device=dev)

        def step():
            worker = multigpu.adaptive_worker(cfg, ctx, rank, world, levels, device_images=device_image)
            if world > 1:
                return multigpu.run_distributed(worker, rank, world)
            return multigpu.run_local([worker])[0]

        def barrier():
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()

        for _ in range(warmup):
            step()
        barrier()
        launches0 = ctx.launch_count()
        t0 = time.perf_counter()
        for _ in range(steps):
            out_levels = step()
        barrier()
        dt = (time.perf_counter() - t0) / steps
        launches = (ctx.launch_count() - launches0) // steps
        out = None
        if rank == 0:
            blocks = [len(L['locs']) for L in out_levels]
            rays = [b * 64 for b in blocks]
            out = {'workload': 'adaptive', 'config': {'workload': WORKLOADS['adaptive'][3] + '; root %dx%d, %d GPU(s)' % (root, root, world)},
                   'value': sum(rays) / dt, 'unit': 'rays/s', 'ms_per_step': 1e3 * dt, 'steps': steps, 'rays_per_level': rays,
                   'blocks_per_level': blocks, 'gpu_launches_rank0_per_step': int(launches),
                   'rank0_stage_ms_last_step': {k: round(1e3 * v, 2) for k, v in multigpu.last_stage_seconds.items()},
                   'image_sum_root': float(np.nansum(out_levels[0]['image'][0])),
                   'timing': 'host wall clock between barriers (camera generation of refined levels, flag all-gather and final '
                             'gather included): end to end by construction'}
        ctx.close()
        return out
    finally:
        shutil.rmtree(workdir, ignore_errors=True)


def finish(res, world, scaling, executed, hbm_peak, fp64_peak):
    """Turn the raw numbers of measure() into the public dict (kernel list with executed rooflines)."""
    st, ms, stages = res.pop('_st'), res.pop('_ms'), res.pop('_stages')
    name, resolution, polarized = res['workload'], res.pop('_resolution'), res.pop('_polarized')
    for k in ('_fp64_peak', '_clocks', '_device', '_passes'):
        res.pop(k, None)
    kernels = kernel_rooflines(name, st, stages, ms['geo'], ms['rad'], fp64_peak, hbm_peak, executed, res['frequencies'], polarized)
    res['config'] = {'workload': workload_label(name, resolution, world, scaling), 'frequencies': res['frequencies'],
                     'rays_rank0': res.pop('rays_rank0'), 'l2': res.pop('l2'),
                     'sharding': 'image rows round-robin over ranks; grid replicated; final gather of the rows to rank 0'}
    res['kernels'] = kernels
    res['dp_attempts_rank0'] = st['num_attempts']
    res['polarized_slab'] = stages['slab']
    res['roofline'] = max(kernels, key=lambda k: k['ms']) if kernels else None
    return res


def main():
    global _JSON_FD
    # stdout must carry exactly one JSON line: everything else that writes to fd 1 (NCCL's version banner,
    # library chatter) is sent to stderr for the whole run
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    args = parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the B200 path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        # a rank that cannot follow (an exception on one side of a collective) must cost minutes, not the default 10
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank), timeout=datetime.timedelta(seconds=240))
    name = args.workload
    scaling = args.scaling
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except (OSError, ValueError):
        pass
    hbm_peak = peaks.get('hbm_gbs', 6650.0)
    executed = load_executed()
    common = {'metric': 'camera rays/sec (geodesic + polarized RT)', 'n_gpus': world, 'warmup': args.warmup, 'higher_is_better': True,
              'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic'}
    try:
        if name == 'adaptive':
            res = measure_adaptive(args.resolution or 512, 3, args.steps, args.warmup, rank, world, local_rank)
            if rank == 0:
                res.update(common)
                res['scaling'] = 'strong'
                emit(res)
            return
        resolution = frame_side(args, name, world)
        res = measure(args, name, resolution, args.steps, args.warmup, rank, world, local_rank, True)
        extras = []
        if not args.no_extras:
            for ex in EXTRAS:
                if ex == name:
                    continue
                try:
                    if ex == 'adaptive':
                        r = measure_adaptive(512, 3, 2, 1, rank, world, local_rank)
                    else:
                        ex_res = (WORKLOADS[ex][2] + world - 1) // world * world
                        r = measure(args, ex, ex_res, 3, 1, rank, world, local_rank, False)
                    if rank == 0:
                        extras.append(r)
                except Exception as e:   # an extra must never cost the headline line
                    if rank == 0:
                        extras.append({'workload': ex, 'error': '%s: %s' % (type(e).__name__, e)})
        if rank == 0:
            fp64_peak, clocks, device = res['_fp64_peak'], res['_clocks'], res['_device']
            units = {'workload': name, 'resolution': resolution, 'n_gpus': world, 'stats_rank0': res['_st'],
                     'passes': res['_passes'], 'frequencies': res['frequencies']}
            res = finish(res, world, scaling, executed, hbm_peak, fp64_peak)
            line = dict(common)
            line.update({'value': res.pop('value'), 'unit': res.pop('unit'), 'steps': args.steps, 'ms_per_step': res.pop('ms_per_step'),
                         'scaling': scaling, 'config': res.pop('config'), 'e2e': res.pop('e2e'),
                         'gpu_launches': res.pop('gpu_launches'), 'roofline': res.pop('roofline')})
            line.update(res)
            line['fp64_peak'] = {'tflops': fp64_peak,
                                 'source': 'DFMA micro-benchmark measured in this run at the clocks recorded below '
                                           '(MEASURED_PEAKS.json has no FP64 entry); the bit-exact geodesic kernel is compiled '
                                           'without FMA and is bounded by half of it'}
            line['hbm_peak'] = {'gbs': hbm_peak, 'source': 'MEASURED_PEAKS.json hbm_gbs (of measured)' if 'hbm_gbs' in peaks
                                else 'fallback 6650 GB/s (of fallback)'}
            line['executed_flops_source'] = {'file': 'profiles/executed_flops.json', 'stale': executed.get('stale'),
                                             'missing': executed.get('missing', False), 'captured': executed.get('captured')}
            line['clocks'] = clocks
            line['device'] = device
            line['timing'] = 'CUDA events on the library stream, barrier + synchronize on both sides, max over ranks'
            ex_out = []
            for r in extras:
                if '_st' in r:
                    r = finish(r, world, 'strong', executed, hbm_peak, fp64_peak)
                    r['steps'] = 3
                ex_out.append(r)
            line['extra'] = ex_out
            if args.dump_units:
                with open(args.dump_units, 'w') as f:
                    json.dump(units, f)
            if world == 1 and not args.no_cpu_baseline:
                threads = os.cpu_count() or 1
                side = args.cpu_resolution or auto_cpu_resolution(name, threads)
                if os.path.exists(REF_BIN):
                    cdir = tempfile.mkdtemp(prefix='bl_cpu_')
                    try:
                        r, c = reference_rays_per_s(make_case(name, cdir, side, 1), threads)
                        line['cpu_baseline'] = {'value': r, 'unit': 'rays/s', 'cores': threads, 'kind': 'reference',
                                                'sample': '%dx%d rays of the same camera and physics, %.1f s of reference compute '
                                                          '(timers geodesic + sample + image)' % (side, side, c)}
                    finally:
                        shutil.rmtree(cdir, ignore_errors=True)
                else:
                    line['cpu_baseline'] = {'value': None, 'unit': 'rays/s', 'cores': threads, 'kind': 'reference',
                                            'sample': 'unavailable: oracle/_ref/blacklight not built'}
            emit(line)
    finally:
        if world > 1:
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
