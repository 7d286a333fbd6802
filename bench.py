#!/usr/bin/env python
"""Benchmark of the per-frame analysis hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA kernels via the C ABI)
  python bench.py --impl reference ...                      reference arm: the reference's OpenCV
                                                            CPU ops on the box's host cores

Headline workload (config.workload): C3 -- 1080p Farneback OpticalFlow + FlowHistogram, the
configuration the metric is quoted on.  A "step" is one pass of the hot path over one batch of
`--pairs` frame pairs (pairs+1 synthetic 1080p RGB frames of a ring larger than L2).
`value`  = frames/s with the frames already resident in HBM (whole job, all ranks).
`e2e`    = frames/s through the host-buffer C-ABI call (stb_pipe_flow): pinned-host frames in,
           H2D copies inside the timed region, 512-byte flow histograms out.
`roofline` = the dominant kernel (fused level-0 update iteration): algorithmic bytes per launch
           / its CUDA-event duration measured in the timed region, against MEASURED_PEAKS.json.
`extra`  = secondary workloads of the metric (4K RGB histogram + shot scoring; C2 640x480 flow).
Frames are independent: ranks own disjoint frame ranges, no collective on the data path
(scaling = weak: per-GPU work is fixed).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic bytes (SURVEY.md §8d / DESIGN.md "Byte model")
FLOW1080_BYTES = 771768000          # canonical Farneback pass model, per 1080p frame
FLOWHIST1080_BYTES = 16589312       # 8*W*H + 512
ITER_BYTES_PER_PX = 80              # level-0 update iteration: read M, R0, R1; write M' (5 f32 each)
HIST4K_BYTES = 24883392             # 3*W*H + 192
FLOW480_BYTES = 114336000


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured'
        except Exception:
            pass
    return 6650.0, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region: the sampler is started
    before the warm-up (nvidia-smi takes a few hundred ms to start) and only rows whose timestamp
    falls inside [t0, t1] of the timed region are kept."""
    Q = ('timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        import datetime
        for seen, r in self.rows:
            f = [x.strip() for x in r.split(',')]
            if len(f) < 9:
                continue
            ts = seen
            try:
                ts = datetime.datetime.strptime(f[0], '%Y/%m/%d %H:%M:%S.%f').timestamp()
            except ValueError:
                pass
            if t0 is not None and not (t0 - 0.02 <= ts <= t1 + 0.02):
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nme, v in zip(names, f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(nme)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max(power) if power else None, 'samples': len(sm), 'reasons': sorted(reasons)}


# ------------------------------------------------------------------------------ reference arm
def _cpu_worker(args):
    kind, h, w, seed, reps = args
    import cv2
    cv2.setNumThreads(1)   # Farneback does not scale with OpenCV threads: one worker per core
    from oracle import cv2_ops
    from scannertools_b200 import synth
    if kind == 'flow':
        clip = synth.textured_clip(seed, 2, h, w)
        t0 = time.perf_counter()
        for _ in range(reps):
            fl = cv2_ops.optical_flow(clip[0], clip[1])
            cv2_ops.flow_histogram(fl)
        return reps, time.perf_counter() - t0
    clip = synth.noise_clip(seed, 2, h, w)
    t0 = time.perf_counter()
    for i in range(reps):
        cv2_ops.histogram(clip[i & 1])
    return reps, time.perf_counter() - t0


def cpu_pass(pool, cores, kind, h, w, reps):
    t0 = time.perf_counter()
    res = pool.map(_cpu_worker, [(kind, h, w, 100 + c, reps) for c in range(cores)])
    wall = time.perf_counter() - t0
    frames = sum(r[0] for r in res)
    return frames, wall


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    """The reference's own CPU implementation of the path: the OpenCV calls its C++ wrappers
    make (oracle/cv2_ops.py), one worker process per host core on disjoint frames."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = host_cores()
    ctx = mp.get_context('spawn')
    with ctx.Pool(cores) as pool:
        cpu_pass(pool, cores, 'flow', 1080, 1920, 1)          # spawn + page-in warm-up
        for _ in range(max(args.warmup - 1, 0)):
            cpu_pass(pool, cores, 'flow', 1080, 1920, 1)
        frames, wall = 0, 0.0
        for _ in range(args.steps):
            f, wl = cpu_pass(pool, cores, 'flow', 1080, 1920, 1)
            frames += f
            wall += wl
    fps = frames / wall
    sample = '%d steps x %d workers x 1 1080p pair each (cv2 %s: cvtColor+Farneback+cartToPolar+calcHist)' % (
        args.steps, cores, __import__('cv2').__version__)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': 'frames/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * wall / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args),
        'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


METRIC = '1080p Farneback OpticalFlow + FlowHistogram throughput'


def workload_config(args):
    return {'workload': 'C3: 1080p (1920x1080) Farneback optical flow (3 levels, winsize 15, 3 iters, polyN 5) + '
                        'FlowHistogram, synthetic textured clip, frame-range sharded',
            'pairs_per_step': args.pairs, 'batch_pairs': args.batch, 'frame': '1920x1080x3 u8',
            'l2_policy': 'inputs larger than L2: ring of %d distinct frames (%.0f MB) + %.0f MB of flow output per step'
                         % (args.pairs + 1, (args.pairs + 1) * 6.2208, args.pairs * 16.5888),
            'parallelism': 'frame-range x%d, no collective' % args.gpus}


# ------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from scannertools_b200 import _lib, ops, synth
    import ctypes as C

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; scannertools_b200 has no CPU fallback '
                         '(use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    # feeder threads and pinned staging buffers on the GPU's own socket (no-op when the box
    # does not expose the topology); STB_NO_NUMA_BIND=1 switches it off for A/B runs
    from scannertools_b200 import sharding
    numa_node = None if os.environ.get('STB_NO_NUMA_BIND') else sharding.bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    lib = _lib.load()
    H, W = 1080, 1920
    P, B = args.pairs, args.batch
    assert P % B == 0, '--pairs must be a multiple of --batch'

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- synthetic clip: this rank's frame range (different content per rank), ring > L2
    clip = synth.textured_clip(1000 + rank, P + 1, H, W)
    host = torch.from_numpy(clip).pin_memory()
    d_frames = host.cuda(non_blocking=True)
    torch.cuda.synchronize()
    of = ops.OpticalFlow(W, H, max_batch=B)
    d_flow = torch.empty((P, H, W, 2), dtype=torch.float32, device='cuda')
    d_fh = torch.empty((P, 2, 64), dtype=torch.int32, device='cuda')
    frame_ptrs = [d_frames[i].data_ptr() for i in range(P + 1)]
    flow_ptrs = [d_flow[i].data_ptr() for i in range(P)]
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)

    def step_device():
        for b0 in range(0, P, B):
            ft = _lib.ptr_table(frame_ptrs[b0:b0 + B + 1])
            ot = _lib.ptr_table(flow_ptrs[b0:b0 + B])
            _lib.check(lib.stb_farneback_run_hist(of._h, ft, B, ot, C.c_void_p(d_fh[b0].data_ptr()), sp), lib)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_device()
    barrier()
    lib.stb_farneback_profile(of._h, 1)
    l0 = lib.stb_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    barrier()
    wall1 = time.time()
    t_dev = ev0.elapsed_time(ev1) * 1e-3
    launches = lib.stb_launch_count() - l0
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    ms_total, nl, npi = C.c_double(), C.c_longlong(), C.c_longlong()
    _lib.check(lib.stb_farneback_profile_read(of._h, C.byref(ms_total), C.byref(nl), C.byref(npi)), lib)
    lib.stb_farneback_profile(of._h, 0)
    fh_dev = d_fh.cpu().numpy().copy()

    # ---- e2e: host frames in (pinned), flow histograms out (pinned), through the host-buffer C ABI.
    # Every step uploads its frames and reads its result back; two calls are kept in flight
    # (stb_pipe_flow_async / stb_pipe_wait: submit step i+1, then wait for and read step i), the way a
    # streaming caller drives it, so one step's uploads hide behind the previous step's kernels.
    pipe = ops.Pipe(W, H, max_batch=B, want_flow=True)
    res = [torch.empty((P, 2, 64), dtype=torch.int32).pin_memory() for _ in range(2)]
    for _ in range(max(1, args.warmup // 2)):
        pipe.wait(pipe.flow_async(host, res[0]))
    fh_e2e = res[0].numpy().copy()
    barrier()
    t0 = time.perf_counter()
    checksum = 0
    prev = pipe.flow_async(host, res[0])
    for i in range(1, args.steps):
        cur = pipe.flow_async(host, res[i & 1])
        pipe.wait(prev)
        checksum += int(res[(i - 1) & 1][0, 0, 0])       # the step's result is read on the host
        prev = cur
    pipe.wait(prev)
    checksum += int(res[(args.steps - 1) & 1][0, 0, 0])
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    barrier()
    assert np.array_equal(fh_e2e, fh_dev), 'e2e and device-resident paths disagree'
    assert np.array_equal(res[(args.steps - 1) & 1].numpy(), fh_dev), 'asynchronous e2e result differs'

    # ---- max over ranks (timing scalars only; no data-path collective)
    times = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device='cuda')
    stats = torch.tensor([float(launches), ms_total.value, float(nl.value), float(npi.value)], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    t_dev, t_e2e = times.tolist()
    launches_all, ms_iter_all, nl_all, npi_all = stats.tolist()

    extra = {}
    if rank == 0 and world == 1 and not args.no_extra:
        extra = extra_workloads(torch, ops, lib, args)
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu_base = cpu_baseline()

    if rank == 0:
        peak, peak_kind = measured_peaks()
        total_frames = world * args.steps * P
        pairs_per_launch = npi_all / nl_all if nl_all else 0.0    # one launch covers a whole batch of pairs
        iter_bytes = ITER_BYTES_PER_PX * H * W * pairs_per_launch
        avg_iter_s = (ms_iter_all / nl_all) * 1e-3 if nl_all else float('nan')
        achieved = iter_bytes / avg_iter_s / 1e9 if nl_all else None
        value = total_frames / t_dev
        line = {
            'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * t_dev / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': dict(workload_config(args), host_numa_node_rank0=numa_node),
            'hbm_roofline_frac_whole_op': (FLOW1080_BYTES + FLOWHIST1080_BYTES) * value / world / (peak * 1e9),
            'roofline': {'bound': 'hbm', 'kernel': 'iter15_tma_kernel<true,false> (level-0 fused box-sum / 2x2 solve / update-matrices iteration, TMA-staged M tiles)',
                         'achieved': achieved, 'peak': peak, 'peak_kind': peak_kind, 'unit': 'GB/s',
                         'frac': (achieved / peak) if achieved else None,
                         'bytes_per_launch': iter_bytes, 'pairs_per_launch': pairs_per_launch, 'avg_launch_us': avg_iter_s * 1e6,
                         'launches_timed': int(nl_all),
                         'share_of_step': (ms_iter_all * 1e-3 / world) / t_dev if t_dev else None,
                         'traffic': NCU_TRAFFIC_BYTES_PER_PAIR * pairs_per_launch,
                         'traffic_source': 'profiles/r01_ncu_full_table.txt (16-pair launch of the same kernel)'},
            'e2e': {'value': total_frames / t_e2e, 'unit': 'frames/s',
                    'h2d_bytes_per_step': world * (P + 1) * H * W * 3 - world * (P // B - 1) * H * W * 3 * 0,
                    'd2h_bytes_per_step': world * P * 512, 'ms_per_step': 1e3 * t_e2e / args.steps,
                    'api': 'stb_pipe_flow_async + stb_pipe_wait (pinned host frames -> flow histograms, two calls in flight)'},
            'gpu_launches': int(launches_all),
            'clocks': clocks,
            'cpu_baseline': cpu_base,
            'extra': extra,
        }
        print(json.dumps(line))
    of.close()
    pipe.close()
    if world > 1:
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum of iter15_tma_kernel<true,false> at level 0 from the
# committed `ncu --set full` capture (profiles/r01_ncu_full_table.txt: grid (40,34,16), 2.204 GB read
# + 644 MB written for the 16 pairs of one launch, mean of its two launches; the reads include the
# L2 prefetch's double fetches, algorithmic bytes are 2.654 GB), per PAIR; scaled by the pairs a bench
# launch processes.
NCU_TRAFFIC_BYTES_PER_PAIR = (2.204e9 + 644.4e6) / 16


def extra_workloads(torch, ops, lib, args):
    """Secondary workloads named by the metric, device-resident + e2e, short loops."""
    import ctypes as C
    from scannertools_b200 import _lib
    peak, _ = measured_peaks()
    out = {}

    def time_dev(fn, iters, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e-3 / iters

    # C4: 4K RGB histogram + shot scores, 32 frames per step (796 MB > L2)
    n4k = 32
    fr = torch.randint(0, 256, (n4k, 2160, 3840, 3), dtype=torch.uint8, device='cuda')
    t = time_dev(lambda: ops.shot_scores(ops.histogram(fr)), 10)
    host4k = fr.cpu().pin_memory()
    pipe = ops.Pipe(3840, 2160, max_batch=7)
    pipe.histogram(host4k)
    t0 = time.perf_counter()
    for _ in range(3):
        pipe.histogram(host4k)
    te = (time.perf_counter() - t0) / 3
    pipe.close()
    out['hist4k'] = {'workload': 'C4: 3840x2160 RGB histogram (16 bins/channel) + shot scores, 32 frames/step, i.i.d. noise',
                     'value': n4k / t, 'unit': 'frames/s', 'ms_per_step': t * 1e3,
                     'roofline': {'bound': 'hbm', 'achieved': n4k * HIST4K_BYTES / t / 1e9, 'peak': peak, 'unit': 'GB/s',
                                  'frac': n4k * HIST4K_BYTES / t / 1e9 / peak},
                     'e2e': {'value': n4k / te, 'unit': 'frames/s', 'h2d_bytes_per_step': n4k * 3840 * 2160 * 3,
                             'd2h_bytes_per_step': n4k * 196}}
    # HSV variant of the shot-detection histogram (old/histograms.py:32-36): fused ConvertToHSV ->
    # Histogram against the two-op chain (ConvertColor writes the HSV frame, Histogram reads it back)
    th = time_dev(lambda: ops.shot_scores(ops.histogram(fr, hsv='COLOR_RGB2HSV')), 10)
    t2 = time_dev(lambda: ops.shot_scores(ops.histogram(ops.convert_color(fr, 'COLOR_RGB2HSV'))), 3, warm=1)
    out['hsvhist4k'] = {'workload': '3840x2160 fused RGB->HSV + histogram (16 bins/channel) + shot scores, 32 frames/step, i.i.d. noise',
                        'value': n4k / th, 'unit': 'frames/s', 'ms_per_step': th * 1e3,
                        'roofline': {'bound': 'hbm', 'achieved': n4k * HIST4K_BYTES / th / 1e9, 'peak': peak, 'unit': 'GB/s',
                                     'frac': n4k * HIST4K_BYTES / th / 1e9 / peak},
                        'unfused_two_op_frames_per_s': n4k / t2}
    del fr, host4k
    # C2: 640x480 Farneback, 16 pairs per step
    from scannertools_b200 import synth
    n480 = 64
    clip = synth.textured_clip(2, n480 + 1, 480, 640)
    d = torch.from_numpy(clip).cuda()
    of = ops.OpticalFlow(640, 480, max_batch=n480)
    o = torch.empty((n480, 480, 640, 2), dtype=torch.float32, device='cuda')
    t = time_dev(lambda: of.execute(d, out=o), 10)
    of.close()
    out['flow480'] = {'workload': 'C2: 640x480 Farneback (3 levels, winsize 15, 3 iters), %d pairs/step' % n480,
                      'value': n480 / t, 'unit': 'frames/s', 'ms_per_step': t * 1e3,
                      'roofline': {'bound': 'hbm', 'achieved': n480 * FLOW480_BYTES / t / 1e9, 'peak': peak, 'unit': 'GB/s',
                                   'frac': n480 * FLOW480_BYTES / t / 1e9 / peak}}
    # C5 in miniature: 8 concurrent 720p streams pinned to this GPU (sharding.stream_assignment puts
    # 8 of the 64 streams on each of 8 GPUs), each with its own OpticalFlow handle (per-stream state
    # never crosses streams), mixed OpticalFlow+FlowHistogram and RGB Histogram work, round-robin
    C5_BYTES = 343008000 + 8 * 1280 * 720 + 512 + 2764992      # flow + flow histogram + RGB histogram per 720p frame
    n_streams, fps_batch = 8, 16
    clips = [torch.from_numpy(synth.textured_clip(300 + sidx, fps_batch + 1, 720, 1280)).cuda() for sidx in range(n_streams)]
    handles = [ops.OpticalFlow(1280, 720, max_batch=fps_batch) for _ in range(n_streams)]

    def c5_step():
        for sidx in range(n_streams):
            handles[sidx].execute_with_histogram(clips[sidx], want_flow=False)
            ops.histogram(clips[sidx][:fps_batch])
    t = time_dev(c5_step, 5, warm=2)
    for hnd in handles:
        hnd.close()
    frames = n_streams * fps_batch
    out['c5_mixed_720p'] = {'workload': 'C5 per-GPU share: %d concurrent 1280x720 streams, OpticalFlow+FlowHistogram and RGB Histogram, '
                                        '%d frames per stream per step' % (n_streams, fps_batch),
                            'value': frames / t, 'unit': 'frames/s', 'ms_per_step': t * 1e3,
                            'roofline': {'bound': 'hbm', 'achieved': frames * C5_BYTES / t / 1e9, 'peak': peak, 'unit': 'GB/s',
                                         'frac': frames * C5_BYTES / t / 1e9 / peak}}
    return out


def cpu_baseline():
    """Reference CPU ops (cv2) on this box's host cores, in a child process that never touches
    CUDA (forking/spawning workers from a process with a live CUDA context is avoided)."""
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), '--cpu-baseline-child'], capture_output=True,
                             text=True, timeout=300)
        for line in out.stdout.splitlines():
            if line.startswith('{'):
                return json.loads(line)
        return {'error': (out.stderr or 'no output')[-300:]}
    except Exception as e:  # noqa: BLE001
        return {'error': repr(e)}


def cpu_baseline_child():
    import multiprocessing as mp
    cores = host_cores()
    ctx = mp.get_context('spawn')
    with ctx.Pool(cores) as pool:
        cpu_pass(pool, cores, 'flow', 1080, 1920, 1)
        reps = 2
        frames, wall = cpu_pass(pool, cores, 'flow', 1080, 1920, reps)
        hframes, hwall = cpu_pass(pool, cores, 'hist', 2160, 3840, 8)
    import cv2
    print(json.dumps({'value': frames / wall, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                      'sample': '%d workers x %d 1080p pairs: cv2 %s cvtColor+Farneback(3,0.5,15,3,5,1.2)+cartToPolar+2xcalcHist, '
                                'cv2.setNumThreads(1) per worker' % (cores, reps, cv2.__version__),
                      'hist4k_frames_per_s': hframes / hwall}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--pairs', type=int, default=64, help='frame pairs per step')
    ap.add_argument('--batch', type=int, default=16, help='pairs per C-ABI batch call')
    ap.add_argument('--no-extra', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--cpu-baseline-child', action='store_true', help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_baseline_child:
        cpu_baseline_child()
        return
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
