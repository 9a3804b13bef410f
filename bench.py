#!/usr/bin/env python
"""Headline benchmark: camera rays/sec of the per-pixel hot path (geodesics + sampling + coefficients +
transfer) on the mock Athena++ snapshot, example_simulation parameters at 1024^2 per GPU.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                     times the unmodified reference's CPU path

A step is one pass of the hot path over one full image of synthetic input.  `value` is rays/s with camera
arrays and grid already resident in HBM; `e2e` is the same metric through the C ABI from pinned HOST
buffers (H2D of the camera arrays and D2H of the image inside the timed region).  One JSON line on rank 0.
"""
import argparse
import json
import math
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

# As-written arithmetic of the reference per unit of geodesic work
FLOP_PER_ATTEMPT = 6900.0
FLOP_PER_ACCEPT = 540.0
FLOP_PER_SAMPLE = 220.0
# As-written arithmetic of the reference per sample of radiation work: sampling 60 flop,
# plasma state + frame geometry 950 flop, and 12 libm calls at the survey's 80 flop-equivalents each
# (acos, atan2, atan, hypot; atan2, atan, sin, cos, 4 hypot); per frequency, thermal unpolarized:
# coefficients 60 flop + transfer 10 flop + 8 libm calls (exp, expm1, cbrt, 2 sqrt, pow; exp, expm1);
# polarized: 6.1 kflop transport/coupling + (thermal 180 flop + 15 calls | kappa 250 flop + 45 calls).
RAD_FLOP_PER_SAMPLE = 60.0 + 950.0 + 12 * 80.0
RAD_FLOP_PER_SAMPLE_FREQ = {'simulation': 70.0 + 8 * 80.0, 'formula': 70.0 + 5 * 80.0, 'polarized': 6100.0 + 250.0 + 45 * 80.0,
                            'polarized_thermal': 6100.0 + 180.0 + 15 * 80.0}
# Algorithmic bytes per sample: trilinear gather of 8 variables and one 64-byte
# step-buffer record written by the geodesic kernel and read once by the radiation kernel
GATHER_BYTES_PER_SAMPLE = 256.0
RECORD_BYTES_PER_SAMPLE = 64.0
# dram__bytes_read.sum + dram__bytes_write.sum per stored sample from the committed ncu --set full captures
# (profiles/r01k_ncu_full_summary.txt): geodesic kernel 65.4 B (all writes), radiation kernel 65.3 B (all reads)
NCU_DRAM_BYTES_PER_SAMPLE = {'geodesic_dp_kernel': 65.4, 'radiate_unpolarized_kernel': 65.3}
# Executed FP64 work from the committed ncu captures (profiles/r01k_ncu_full_summary.txt): thread-level
# DADD + DMUL + 2 DFMA per unit (DP attempt for the geodesic kernel, stored sample for the radiation kernels, at the
# workload's frequency count) and the share of cycles the FP64 pipe was active.  Unlike the as-written counts above
# (the reference's operation count, the reproducible contract of SURVEY 8d) these are what the restructured kernels
# really issue; a kernel keyed by workload has no entry for workloads that were not captured.
NCU_EXECUTED = {
    'geodesic_dp_kernel': {'flop_per_unit': 1.577e11 / 19384738, 'unit': 'attempt', 'fp64_pipe_active': 0.670},
    'simulation': {'flop_per_unit': 1.544e11 / 184675182, 'unit': 'sample', 'fp64_pipe_active': 0.301, 'dram_bytes_per_sample': 65.3},
    'polarized_thermal': {'flop_per_unit': 4.495e11 / 103872314, 'unit': 'sample', 'fp64_pipe_active': 0.277, 'dram_bytes_per_sample': 67.5},
    'polarized': {'flop_per_unit': 5.414e11 / 46164658, 'unit': 'sample (4 frequencies)', 'fp64_pipe_active': 0.294, 'dram_bytes_per_sample': 69.3},
}


def executed_roofline(key, units, ms, fp64_peak):
    e = NCU_EXECUTED.get(key)
    if e is None or ms <= 0:
        return None
    tflops = e['flop_per_unit'] * units / (ms * 1e-3) / 1e12
    return {'achieved': tflops, 'unit': 'TFLOP/s', 'frac': tflops / fp64_peak if fp64_peak else None,
            'flop_per_unit': e['flop_per_unit'], 'per': e['unit'], 'fp64_pipe_active': e['fp64_pipe_active'],
            'source': 'profiles/r01k_ncu_full_summary.txt (ncu DADD + DMUL + 2 DFMA thread instructions per unit x live unit count / live kernel time)'}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--resolution', type=int, default=1024, help='image side per GPU (weak scaling)')
    ap.add_argument('--workload', default='simulation', choices=['simulation', 'formula', 'polarized', 'polarized_thermal'])
    ap.add_argument('--tile-rays', type=int, default=0)
    ap.add_argument('--grid-scale', type=int, default=1,
                    help='refine the mock snapshot by this factor per dimension (4: 308x256x512 cells, 1.3 GB of primitives -- '
                         'the gather leaves L2 and becomes HBM traffic; SURVEY.md section 8d)')
    ap.add_argument('--cpu-resolution', type=int, default=0, help='side of the bounded CPU sample (0 = auto)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    return ap.parse_args()


def workload_case(args, workdir, resolution, write_mock):
    from harness import Case
    base = {'simulation': 'simulation.input', 'formula': 'formula.input', 'polarized': 'simulation.input',
            'polarized_thermal': 'simulation.input'}[args.workload]
    over = {'camera_resolution': resolution}
    if args.workload == 'polarized':
        over.update({'image_polarization': 'true', 'image_num_frequencies': 4, 'image_frequency_start': '8.6e10',
                     'image_frequency_end': '3.45e11', 'image_frequency_spacing': 'log', 'plasma_kappa_frac': '1.0',
                     'plasma_kappa': '4.0', 'plasma_w': '1.0'})
    mock = None
    if args.grid_scale > 1 and base == 'simulation.input':
        k = args.grid_scale
        mock = {'n_r': 77 * k, 'n_th': 64 * k, 'n_ph': 128 * k}
    if args.workload == 'polarized_thermal':
        over.update({'image_polarization': 'true'})
    # host threads (camera pixels, reader conversions) per rank: the box's cores shared among the ranks
    world = int(os.environ.get('WORLD_SIZE', '1'))
    case = Case(workdir, base, over, mock=mock, threads=max(1, (os.cpu_count() or 1) // world))
    return case


def workload_name(args, res_total, n_gpus):
    g = '%dx%dx%d' % (77 * args.grid_scale, 64 * args.grid_scale, 128 * args.grid_scale)
    d = {'simulation': 'mock Athena++ snapshot (' + g + ' SKS, generate_mock_simulation defaults), example_simulation '
                       'parameters: unpolarized thermal synchrotron, trilinear sampling, DP geodesics',
         'formula': 'example_formula parameters: formula plasma, DP geodesics',
         'polarized': 'mock Athena++ snapshot, polarized kappa=4 synchrotron, 4 frequencies',
         'polarized_thermal': 'mock Athena++ snapshot, polarized thermal synchrotron, 1 frequency'}[args.workload]
    return '%s; image %dx%d over %d GPU(s)' % (d, res_total, res_total, n_gpus)


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


def reference_rays_per_s(case, threads):
    """Run the unmodified reference once on `case`; rays/s from its own timers (geodesic + sample + image)."""
    case.kv['num_threads'] = str(threads)
    t0 = time.time()
    ref = case.run_reference(checkpoints=False)
    wall = time.time() - t0
    t = ref['timers']
    compute = t.get('Integrating geodesics', 0.0) + t.get('Sampling simulation', 0.0) + t.get('Integrating image', 0.0) \
        + t.get('Rendering', 0.0)
    rays = int(case.kv['camera_resolution']) ** 2
    return rays / compute, compute, wall


def auto_cpu_resolution(args, threads):
    # ~3k rays/s on 8 threads measured in the survey container; aim for ~15 s of CPU work
    est = 360.0 * threads * (0.5 if args.workload == 'formula' else 1.0) * (0.1 if args.workload == 'polarized' else 1.0)
    side = int(math.sqrt(est * 15.0))
    return max(32, min(256, side // 8 * 8))


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    from harness import REF_BIN
    threads = os.cpu_count() or 1
    line = {'impl': 'reference', 'metric': 'camera rays/sec (geodesic+RT)', 'unit': 'rays/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic'}
    if not os.path.exists(REF_BIN):
        emit({'impl': 'reference', 'unavailable': 'oracle/_ref/blacklight was not built (no /root/reference at build time)'})
        return
    side = args.cpu_resolution or auto_cpu_resolution(args, threads)
    workdir = tempfile.mkdtemp(prefix='bl_ref_')
    try:
        case = workload_case(args, workdir, side, True)
        rates, computes = [], []
        for i in range(args.warmup + args.steps):
            r, c, _ = reference_rays_per_s(case, threads)
            if i >= args.warmup:
                rates.append(r)
                computes.append(c)
            if i == 0 and c > 60.0:   # keep the whole run within a few minutes
                args.warmup = 0
                args.steps = max(1, min(args.steps, int(120.0 / c)))
                rates, computes = [r], [c]
                if args.steps == 1:
                    break
        rays = side * side
        value = rays * len(computes) / sum(computes)
        res_total = int(round(args.resolution * math.sqrt(args.gpus)))
        line.update({'value': value, 'ms_per_step': 1e3 * sum(computes) / len(computes), 'steps': len(computes),
                     'warmup': args.warmup,
                     'config': {'workload': workload_name(args, res_total, args.gpus),
                                'note': 'reference CPU path timed on a bounded sample of the same camera and physics'},
                     'cpu_baseline': {'value': value, 'unit': 'rays/s', 'cores': threads, 'kind': 'reference',
                                      'sample': '%dx%d rays of the same camera (rays/s is resolution independent); '
                                                'reference timers geodesic+sample+image' % (side, side)},
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
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import blacklight_b200 as bl
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the B200 path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's version/debug banner goes to stderr
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    n_gpus = world

    # weak scaling: the image grows so that every GPU keeps resolution^2 rays; rows are dealt round-robin
    res_total = args.resolution if n_gpus == 1 else int(math.ceil(args.resolution * math.sqrt(n_gpus) / (8 * n_gpus))) * 8 * n_gpus
    workdir = tempfile.mkdtemp(prefix='bl_bench_%d_' % rank)
    try:
        case = workload_case(args, workdir, res_total, False)
        cfg = case.config(device=local_rank, tile_rays=args.tile_rays)
        ctx = bl.Context(cfg)
        info = ctx.device_info()
        if case.sim:
            ctx.upload_grid(case.grid_arrays())
        # this rank's rays: image rows rank, rank+world, ... (cost varies strongly across the image)
        pos_all, dir_all, fac_all = cfg.camera_root()
        rows = np.arange(rank, res_total, n_gpus)
        idx = (rows[:, None] * res_total + np.arange(res_total)[None, :]).ravel()
        n_rays = len(idx)
        pos = torch.from_numpy(pos_all[idx]).pin_memory()
        dirs = torch.from_numpy(dir_all[idx]).pin_memory()
        fac = torch.from_numpy(fac_all[idx]).pin_memory()
        del pos_all, dir_all, fac_all
        Q = ctx.num_quantities
        image_host = torch.empty((Q, n_rays), dtype=torch.float64).pin_memory()
        image_np = image_host.numpy()
        pos_np, dir_np, fac_np = pos.numpy(), dirs.numpy(), fac.numpy()
        gathered = None
        if world > 1:
            image_dev = torch.empty((Q, n_rays), dtype=torch.float64, device='cuda')
            gathered = [torch.empty_like(image_dev) for _ in range(world)] if rank == 0 else None

        # every kernel and copy of the library is issued on the context's own stream: time with CUDA events
        # recorded on THAT stream (torch's current stream sees none of it)
        lib_stream = torch.cuda.ExternalStream(ctx.cuda_stream(), device=torch.device('cuda', local_rank))

        def barrier():
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        def timed(fn, steps):
            """barrier + sync, `steps` calls of fn bracketed by events on the library stream, sync + barrier;
            returns (device seconds, host seconds, last result)."""
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            t0 = time.perf_counter()
            ev0.record(lib_stream)
            out = None
            for _ in range(steps):
                out = fn()
            lib_stream.wait_stream(torch.cuda.current_stream())   # the NCCL gather of the last step (N > 1)
            ev1.record(lib_stream)
            ev1.synchronize()
            barrier()
            return ev0.elapsed_time(ev1) * 1e-3, time.perf_counter() - t0, out

        def step_e2e():
            st0 = ctx.trace_level(0, pos_np, dir_np, fac_np)            # H2D of camera arrays (+ trace if resident)
            _, _, st = ctx.radiate_level(0, image=image_np)               # kernels + D2H of the image
            if world > 1:                                                 # final image gather over NVLink
                image_dev.copy_(image_host, non_blocking=True)
                dist.gather(image_dev, gathered, dst=0)
            return st

        def step_resident():
            ctx.retrace_level(0)
            _, _, st = ctx.radiate_level(0, download=False)
            return st

        for _ in range(args.warmup):
            step_e2e()
        fp64_peak = ctx.measure_fp64_peak()

        sampler = ClockSampler(local_rank)
        # ---- end-to-end from host buffers ----
        t_e2e, wall_e2e, _ = timed(step_e2e, args.steps)
        # ---- resident: camera arrays and grid already in HBM, no image download ----
        launches0 = ctx.launch_count()
        ms_acc = {'geo': 0.0, 'rad': 0.0}

        def step_resident_acc():
            st_ = step_resident()
            ms_acc['geo'] += st_['ms_geodesic']
            ms_acc['rad'] += st_['ms_radiation']
            return st_

        t_res, wall_res, st = timed(step_resident_acc, args.steps)
        ms_geo, ms_rad = ms_acc['geo'], ms_acc['rad']
        launches = ctx.launch_count() - launches0
        clocks = sampler.stop()

        times = torch.tensor([t_e2e, t_res], dtype=torch.float64, device='cuda')
        counts = torch.tensor([float(n_rays), float(launches)], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(times, op=dist.ReduceOp.MAX)
            dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        t_e2e, t_res = times.tolist()
        total_rays, total_launches = counts.tolist()

        if rank == 0:
            K = args.steps
            F = int(cfg.keys.get('image_num_frequencies', '1'))
            value = total_rays * K / t_res
            e2e = total_rays * K / t_e2e
            # rooflines of both kernels on this rank; `roofline` carries the dominant (slower) one
            geo_ms, rad_ms = ms_geo / K, ms_rad / K
            flop = st['num_attempts'] * FLOP_PER_ATTEMPT + st['num_accepted'] * FLOP_PER_ACCEPT + st['num_samples'] * FLOP_PER_SAMPLE
            geo_tflops = flop / (geo_ms * 1e-3) / 1e12 if geo_ms > 0 else 0.0
            rad_flop = st['num_samples'] * (RAD_FLOP_PER_SAMPLE + F * RAD_FLOP_PER_SAMPLE_FREQ[args.workload])
            rad_tflops = rad_flop / (rad_ms * 1e-3) / 1e12 if rad_ms > 0 else 0.0
            gather_gbs = st['num_samples'] * GATHER_BYTES_PER_SAMPLE / (rad_ms * 1e-3) / 1e9 if rad_ms > 0 else 0.0
            peaks = {}
            try:
                peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
            except (OSError, ValueError):
                pass
            hbm_peak = peaks.get('hbm_gbs', 6650.0)
            hbm_src = 'MEASURED_PEAKS.json hbm_gbs (of measured)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s (of fallback)'
            peak_src = 'DFMA micro-benchmark measured in this run (MEASURED_PEAKS.json has no FP64 entry); no-FMA code such as the bit-exact geodesic kernel is bounded by half of it'
            rad_name = 'radiate_polarized_kernel' if args.workload.startswith('polarized') else 'radiate_unpolarized_kernel'
            roofs = {
                'geodesic_dp_kernel': {'kernel': 'geodesic_dp_kernel', 'bound': 'fp64', 'achieved': geo_tflops, 'peak': fp64_peak,
                                       'unit': 'TFLOP/s', 'frac': geo_tflops / fp64_peak if fp64_peak else None,
                                       'traffic': st['num_samples'] * NCU_DRAM_BYTES_PER_SAMPLE['geodesic_dp_kernel'],
                                       'ms': geo_ms, 'peak_source': peak_src,
                                       'executed': executed_roofline('geodesic_dp_kernel', st['num_attempts'], geo_ms, fp64_peak),
                                       'hbm': {'achieved': st['num_samples'] * RECORD_BYTES_PER_SAMPLE / (geo_ms * 1e-3) / 1e9 if geo_ms > 0 else 0.0,
                                               'peak': hbm_peak, 'unit': 'GB/s', 'peak_source': hbm_src}},
                rad_name: {'kernel': rad_name, 'bound': 'fp64', 'achieved': rad_tflops, 'peak': fp64_peak, 'unit': 'TFLOP/s',
                           'frac': rad_tflops / fp64_peak if fp64_peak else None,
                           'traffic': (st['num_samples'] * NCU_EXECUTED[args.workload]['dram_bytes_per_sample']
                                       if args.workload in NCU_EXECUTED and args.grid_scale == 1 else None),
                           'ms': rad_ms, 'peak_source': peak_src,
                           'executed': (executed_roofline(args.workload, st['num_samples'], rad_ms, fp64_peak)
                                        if F == (4 if args.workload == 'polarized' else 1) else None),
                           'note': 'as-written flop-equivalents of the reference per sample (libm calls at 80); the cell gather is '
                                   'L2 resident for the 20 MB mock grid',
                           'gather': {'achieved': gather_gbs, 'unit': 'GB/s', 'bytes_per_sample': GATHER_BYTES_PER_SAMPLE},
                           'hbm': {'achieved': st['num_samples'] * RECORD_BYTES_PER_SAMPLE / (rad_ms * 1e-3) / 1e9 if rad_ms > 0 else 0.0,
                                   'peak': hbm_peak, 'unit': 'GB/s', 'peak_source': hbm_src}},
            }
            roofline = roofs['geodesic_dp_kernel'] if geo_ms >= rad_ms else roofs[rad_name]
            other = roofs[rad_name] if geo_ms >= rad_ms else roofs['geodesic_dp_kernel']
            line = {
                'metric': 'camera rays/sec (geodesic+RT)', 'value': value, 'unit': 'rays/s', 'n_gpus': n_gpus, 'steps': K,
                'warmup': args.warmup, 'ms_per_step': 1e3 * t_res / K, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
                'config': {'workload': workload_name(args, res_total, n_gpus), 'rays_per_gpu': n_rays, 'frequencies': F,
                           'l2': 'inputs larger than L2: %.1f GB step buffer written and re-read per step' %
                                 (st['num_samples'] * 64 / 1e9),
                           'sharding': 'image rows round-robin over ranks; grid replicated; final image gather'},
                'e2e': {'value': e2e, 'unit': 'rays/s', 'ms_per_step': 1e3 * t_e2e / K,
                        'h2d_bytes_per_step': int(total_rays * 72), 'd2h_bytes_per_step': int(total_rays * 8 * Q)},
                'gpu_launches': int(total_launches),
                'kernels': {'geodesic_ms_per_step': geo_ms, 'radiation_ms_per_step': rad_ms,
                            'samples_per_step': st['num_samples'], 'dp_attempts_per_step': st['num_attempts'],
                            'geodesic_tflops_as_written': geo_tflops, 'radiation_tflops_as_written': rad_tflops,
                            'fp64_peak_tflops_measured': fp64_peak,
                            'radiation_gather_gbs': gather_gbs, 'ray_freq_per_s': value * F,
                            'host_wall_ms_per_step': 1e3 * wall_res / K, 'host_wall_ms_per_step_e2e': 1e3 * wall_e2e / K,
                            'polarized_stages_last_step': ctx.polarized_stage_ms(0)},
                'roofline': roofline, 'roofline_other_kernel': other, 'clocks': clocks, 'device': info['name'],
                'timing': 'CUDA events on the library stream, barrier + synchronize on both sides, max over ranks',
            }
            if n_gpus == 1 and not args.no_cpu_baseline:
                from harness import REF_BIN
                threads = os.cpu_count() or 1
                side = args.cpu_resolution or auto_cpu_resolution(args, threads)
                if os.path.exists(REF_BIN):
                    cdir = tempfile.mkdtemp(prefix='bl_cpu_')
                    try:
                        ccase = workload_case(args, cdir, side, True)
                        r, c, wall = reference_rays_per_s(ccase, threads)
                        line['cpu_baseline'] = {'value': r, 'unit': 'rays/s', 'cores': threads, 'kind': 'reference',
                                                'sample': '%dx%d rays of the same camera, %.1f s of reference compute '
                                                          '(timers geodesic+sample+image)' % (side, side, c)}
                    finally:
                        shutil.rmtree(cdir, ignore_errors=True)
                else:
                    line['cpu_baseline'] = {'value': None, 'unit': 'rays/s', 'cores': threads, 'kind': 'reference',
                                            'sample': 'unavailable: oracle/_ref/blacklight not built'}
            emit(line)
        ctx.close()
    finally:
        shutil.rmtree(workdir, ignore_errors=True)
        if world > 1:
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
