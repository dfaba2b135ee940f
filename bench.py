#!/usr/bin/env python
"""bench.py -- frames/s of MobilePoser's per-frame hot path on synthetic 5-IMU windows.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg2|cfg4] [--min-seconds S]

Workloads (BASELINE.json configs):
  cfg3 (default): MobilePoserNet 4 heads + kinematic tail (K5) + translation (K6) + the PHYSICS-hook optimizer (K8,
                  --physics off to leave it out like the reference's default PHYSICS=0) = forward_offline on a batch
                  of 256 sequences x 300 frames per GPU (weak scaling: every rank gets its own 256 sequences);
  cfg2:           the same path at batch 1 (one 300-frame window), also reported inside the cfg3 line as "batch1".
  cfg4:           the evaluate.py-shaped synthetic DIP set (50 sequences x 3000 frames) through evaluate_pose, sequences sharded over
                  the GPUs (strong scaling), one all-gather of the [50, 8, 2] metric rows per pass; also inside the cfg3 line as "cfg4".
The cfg3 line also carries "pinned_path": the same workload with the physics hook off (the reference's default PHYSICS=0, the
path whose parity is pinned to the reference), with its own pipelined value and e2e; the CPU arm reports the matching number.
A "step" is one forward_offline pass over the batch.  `value` = frames/s with inputs resident in HBM,
`e2e` = the same through the host-buffer C-ABI entry (pinned host imu in, pose/joints/tran/contact out),
`roofline` = the dominant kernel (cluster LSTM recurrence, H=256) against the measured HBM peak,
`cpu_baseline` = the oracle port (the reference's torch-CPU path) timed on this box's host cores.

--impl reference times the reference's CPU implementation of the path (oracle/torch_port.py, which calls the
same torch CPU nn.LSTM/Linear kernels the reference calls) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# The CPU arm (--impl reference) uses every host core.  torchrun exports OMP_NUM_THREADS=1 to its children, and the
# OpenMP / oneDNN pools read it when torch is imported: fix the environment and re-exec before that happens.
if 'reference' in sys.argv and os.environ.get('MP_BENCH_REEXEC') != '1':
    _n = str(os.cpu_count() or 1)
    if os.environ.get('OMP_NUM_THREADS', _n) != _n or os.environ.get('MKL_NUM_THREADS', _n) != _n:
        os.environ.update(OMP_NUM_THREADS=_n, MKL_NUM_THREADS=_n, MP_BENCH_REEXEC='1')
        os.execv(sys.executable, [sys.executable] + sys.argv)

import torch

T_FRAMES = 300
METRIC = 'frames/sec (batch=1 and 256) synthetic 5-IMU@60Hz at 1/2/4/8 B200 vs CPU ref'
IO_BYTES_PER_FRAME = 240 + 864 + 288 + 12 + 8          # imu in; pose, joints, tran, contact out
WEIGHT_BYTES = 26_699_976


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg3', choices=['cfg3', 'cfg2', 'cfg4'])
    ap.add_argument('--batch', type=int, default=0, help='override sequences per GPU')
    ap.add_argument('--cpu-sample', type=int, default=256, help='sequences of the workload the CPU arm times per step (256 = all of cfg3)')
    ap.add_argument('--min-seconds', type=float, default=3.0,
                    help='sustained measurement: after the contract region of exactly K steps, the same K steps are repeated until the '
                         'timed region is at least this long, and `value` is taken over all of them (0 = the K steps only)')
    ap.add_argument('--no-cfg4', action='store_true', help='skip the sharded evaluate_pose (cfg4) section')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--depth', type=int, default=0, help='batches in flight per GPU (pipeline over steps); default 6 for cfg3, 1 for the latency configuration cfg2')
    ap.add_argument('--physics', default='auto', choices=['auto', 'on', 'off'],
                    help='K8 optimizer behind the PHYSICS hook (auto: on for cfg3, which names it; off for cfg2)')
    return ap.parse_args()


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def __enter__(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '200', '-i', str(self.gpu_index)],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if not self.path or not os.path.exists(self.path):
            return out
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            p = [v.strip() for v in line.split(',')]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), p[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def dist_env():
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    return rank, world, local


# ------------------------------------------------------------------------------------------------------
def physics_on(args):
    return args.physics == 'on' or (args.physics == 'auto' and args.workload == 'cfg3')


def use_all_host_threads():
    """The CPU arm uses every host core; torchrun exports OMP_NUM_THREADS=1 to its children, undo that here."""
    n = os.cpu_count() or 1
    torch.set_num_threads(n)
    from oracle.physics_c import set_threads
    set_threads(n)
    return n


def cpu_reference_pass(oracle, x, lens, physics=False):
    """The reference's CPU path for a batch: batched net.forward (net.py:101-119) + the per-sequence
    translation tail of forward_offline (net.py:125-154) + (physics) the PHYSICS-hook loop of net.py:157-169 as the
    compiled float64 statement of K8 (oracle/physics_port.c, one OpenMP thread per sequence)."""
    from oracle.torch_port import offline_translation
    oracle.vel_state = None
    pose, joints, vel, contact = oracle.forward(x, lens)
    B, T = x.shape[0], x.shape[1]
    vel = vel.view(B, T, 72)
    trans = [offline_translation(joints[b, :L], vel[b, :L], contact[b, :L]) for b, L in enumerate(lens)]
    if physics:
        from oracle.physics_c import PhysicsOptimizerC
        pose_opt, _ = PhysicsOptimizerC(B=B).optimize_sequences(pose.view(B, T, 24, 3, 3).numpy(), vel.numpy(), contact.numpy(), lens)
        pose = torch.from_numpy(pose_opt)
    return pose, joints, trans, contact


def time_cpu_sample(sd, x_sample, budget_s=20.0, min_passes=2, physics=False):
    """frames/s of the oracle port on this box's host cores for a bounded sample; returns the cpu_baseline dict."""
    from oracle.torch_port import OraclePoser
    use_all_host_threads()
    oracle = OraclePoser(sd)
    B, T = x_sample.shape[0], x_sample.shape[1]
    lens = [T] * B
    with torch.no_grad():
        cpu_reference_pass(oracle, x_sample[:2], lens[:2], physics)      # warm-up (MKL thread pools)
        t0 = time.perf_counter()
        passes = 0
        while passes < min_passes or (time.perf_counter() - t0 < budget_s * 0.6 and passes < 50):
            cpu_reference_pass(oracle, x_sample, lens, physics)
            passes += 1
        dt_batched = (time.perf_counter() - t0) / passes
        # the reference's own evaluate.py call pattern: one sequence at a time (forward_offline, B = 1)
        t1 = time.perf_counter()
        n1 = 0
        while n1 < 2 or (time.perf_counter() - t1 < budget_s * 0.3 and n1 < B):
            oracle.vel_state = None
            if physics:
                cpu_reference_pass(oracle, x_sample[n1 % B:n1 % B + 1], [T], True)
            else:
                oracle.forward_offline(x_sample[n1 % B:n1 % B + 1], [T])
            n1 += 1
        dt_single = (time.perf_counter() - t1) / n1
    fps_batched, fps_single = B * T / dt_batched, T / dt_single
    return {
        'value': max(fps_batched, fps_single), 'unit': 'frames/s', 'cores': torch.get_num_threads(),
        'host_cpus': os.cpu_count(), 'kind': 'port',
        'sample': (f'{B} of the workload\'s sequences x {T} frames: batched forward + per-sequence translation tail '
                   + ('+ K8 (C float64 port, OpenMP) ' if physics else '') +
                   f'= {fps_batched:.0f} frames/s ({passes} passes); evaluate.py-style one sequence at a time '
                   f'(forward_offline, B=1) = {fps_single:.0f} frames/s ({n1} sequences); torch {torch.__version__} CPU'),
        'batched_fps': fps_batched, 'single_sequence_fps': fps_single,
    }


def cpu_small_configs(sd, budget_s=3.0):
    """The CPU-runnable configurations of BASELINE.json on this box's host cores (oracle port): cfg1 (joints head only, one
    300-frame sequence -- overfit.py's shape), cfg2 (full forward_offline, batch 1) and cfg5 (forward_online ticks, window 45).
    Batch-1 work is often fastest single-threaded: both thread counts are timed, the better one is reported."""
    from mobileposer_b200.synthetic import synthetic_imu
    from oracle.torch_port import OraclePoser
    oracle = OraclePoser(sd)
    x = synthetic_imu(90000, T_FRAMES)[None]
    n_all = os.cpu_count() or 1

    def best(fn, frames):
        out = {}
        for nt in sorted({1, n_all}):
            torch.set_num_threads(nt)
            fn()
            t0, n = time.perf_counter(), 0
            while n < 2 or time.perf_counter() - t0 < budget_s / 2:
                fn()
                n += 1
            out[nt] = (time.perf_counter() - t0) / n
        nt = min(out, key=out.get)
        torch.set_num_threads(n_all)
        return {'value': frames / out[nt], 'unit': 'frames/s', 'ms_per_call': out[nt] * 1e3, 'threads': nt,
                'ms_per_call_by_threads': {str(k): v * 1e3 for k, v in out.items()}}

    def cfg2():
        oracle.vel_state = None
        oracle.forward_offline(x, [T_FRAMES])

    def cfg5():
        oracle.forward_online(x[0, 7])

    with torch.no_grad():
        res = {'cfg1_joints_only_T300': best(lambda: oracle.joints_head(x, [T_FRAMES]), T_FRAMES),
               'cfg2_forward_offline_B1_T300': best(cfg2, T_FRAMES)}
        oracle.vel_state = None
        oracle.reset_online()
        res['cfg5_forward_online_tick_W45'] = best(cfg5, 1)
        res['cfg5_forward_online_tick_W45']['note'] = 'one stream; the reference recomputes the whole 45-frame window per tick (net.py:173-219)'
    oracle.vel_state = None
    return res


def seeded_state_dict():
    import mobileposer_b200 as mp
    torch.manual_seed(0)
    net = mp.MobilePoserNet().eval()
    return net, {k: v.clone() for k, v in net.state_dict().items()}


# ------------------------------------------------------------------------------------------------------
def bench_config(args, B, phys):
    """The `config` object of the JSON line -- the same for both arms (the driver compares them)."""
    return {'workload': workload_name(args.workload, B, phys), 'combo': 'lw_rp', 'frames_per_sequence': T_FRAMES,
            'weights': 'torch.manual_seed(0) default init (bit-identical to the reference init)'}


def run_reference(args):
    rank, world, local = dist_env()
    if rank != 0:
        return
    from mobileposer_b200.synthetic import synthetic_imu_batch
    from oracle.torch_port import OraclePoser
    n_threads = use_all_host_threads()
    if args.workload == 'cfg4':
        return run_reference_cfg4(args, n_threads)
    B = (args.batch or (256 if args.workload == 'cfg3' else 1))
    Bs = min(B, args.cpu_sample)
    _, sd = seeded_state_dict()
    x = synthetic_imu_batch(list(range(Bs)), T_FRAMES)
    oracle = OraclePoser(sd)
    lens = [T_FRAMES] * Bs
    phys = physics_on(args)

    def timed(physics, steps, warmup):
        with torch.no_grad():
            for _ in range(warmup):
                cpu_reference_pass(oracle, x, lens, physics)
            t0 = time.perf_counter()
            for _ in range(steps):
                cpu_reference_pass(oracle, x, lens, physics)
            return time.perf_counter() - t0

    dt = timed(phys, args.steps, args.warmup)
    fps = Bs * T_FRAMES * args.steps / dt
    what = (f'{Bs} of {B} sequences x {T_FRAMES} frames per step' if Bs < B else f'all {B} sequences x {T_FRAMES} frames per step')
    sample = (what + ', batched forward + per-sequence translation tail' + (' + K8 (C float64 port, OpenMP)' if phys else '') +
              f', torch {torch.__version__} CPU, {n_threads} threads')
    pinned = None
    if phys:        # the path whose parity is pinned (the reference's default PHYSICS=0), a few steps
        k0 = max(2, min(args.steps, 5))
        dt0 = timed(False, k0, 1)
        pinned = {'value': Bs * T_FRAMES * k0 / dt0, 'unit': 'frames/s', 'ms_per_step': dt0 / k0 * 1e3, 'steps': k0,
                  'workload': workload_name(args.workload, B, False)}
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': 'frames/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': bench_config(args, B, phys),
        'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': n_threads, 'host_cpus': os.cpu_count(),
                         'kind': 'port', 'sample': sample},
        'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'pinned_path': pinned, 'small_configs': cpu_small_configs(sd), 'gpu_launches': 0,
    }
    emit(json.dumps(line))


def run_reference_cfg4(args, n_threads):
    """cfg4 on the host cores: evaluate.py's loop (reset + forward_offline per sequence, metric rows) over a bounded sample of
    the synthetic DIP set, through the oracle port."""
    from mobileposer_b200.evaluate import PoseEvaluator, r6d_to_rotation_matrix, synthetic_dip
    from oracle.torch_port import OraclePoser
    _, sd = seeded_state_dict()
    oracle = OraclePoser(sd)
    n_seq = max(1, min(50, args.cpu_sample if args.cpu_sample < 256 else 4))
    items = synthetic_dip(n_subjects=1, n_seq=n_seq)
    ev = PoseEvaluator()

    def one_pass():
        for imu, pose_t, _, tran_t in items:
            oracle.reset()
            pose, _, tran, _ = oracle.forward_offline(imu[None], [imu.shape[0]])
            ev.eval(pose, r6d_to_rotation_matrix(pose_t).view(-1, 24, 3, 3), tran_p=tran, tran_t=tran_t)

    with torch.no_grad():
        for _ in range(min(args.warmup, 1)):
            one_pass()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            one_pass()
        dt = time.perf_counter() - t0
    frames = sum(it[0].shape[0] for it in items)
    fps = frames * args.steps / dt
    sample = (f'{n_seq} of the 50 sequences x 3000 frames per step: evaluate.py loop (forward_offline B=1 + metric rows), '
              f'torch {torch.__version__} CPU, {n_threads} threads')
    emit(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': 'frames/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': bench_config(args, 50, False),
        'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': n_threads, 'host_cpus': os.cpu_count(), 'kind': 'port', 'sample': sample},
        'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}))


def workload_name(w, B, physics=False):
    if w == 'cfg4':
        return ('cfg4: evaluate.py-shaped synthetic DIP (10 subjects x 5 sequences x 3000 frames, combo lw_rp): evaluate_pose = '
                'forward_offline + metric rows per sequence, sequences sharded over the GPUs, one all-gather of [50, 8, 2] rows')
    if w == 'cfg3':
        return (f'cfg3: MobilePoserNet 4 heads + kinematic tail + translation (forward_offline)'
                + (' + physics-hook optimizer K8 (parity unpinned: the reference module is absent)' if physics else '')
                + f', batch={B} sequences x {T_FRAMES} frames per GPU, combo lw_rp')
    return f'cfg2: full MobilePoserNet forward_offline, batch={B}, {T_FRAMES}-frame window, combo lw_rp'


def timed_device_steps(fn, steps, warmup, dist, streams=(), min_seconds=0.0):
    """W warm-ups, then K steps between CUDA events, barrier + synchronize on both sides; max over ranks.
    `streams`: the streams the steps are enqueued on when that is not the current one -- they start after the first
    event and the second event waits for all of them.  With `min_seconds` the K-step region is followed by ONE longer region
    of R x K steps (R chosen so that it lasts at least that long) and that one is returned: -> (ms, steps_timed)."""
    for i in range(warmup):
        fn(i)

    def region(n, off):
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for st in streams:
            st.wait_event(e0)
        for i in range(n):
            fn(off + i)
        for st in streams:
            torch.cuda.current_stream().wait_stream(st)
        e1.record()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    ms = region(steps, warmup)
    n = steps
    if min_seconds > 0 and ms < min_seconds * 1e3:
        reps = int(min(1000, -(-min_seconds * 1e3 // max(ms, 1e-3))))
        if dist is not None:                        # every rank must run the same number of steps
            t = torch.tensor([reps], device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            reps = int(t.item())
        n = reps * steps
        ms = region(n, warmup + steps)
    return ms, n


def bind_rank_to_cores(local, world):
    """One slice of the host cores per rank (the e2e path is a Python submit / wait loop plus DMA into pinned host memory:
    eight ranks floating over the same cores cost the N = 8 run a quarter of its end-to-end rate in round 1).  Pinned buffers
    are allocated after this, so first-touch places them next to the cores that use them."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        if world <= 1 or len(cores) < 2 * world:
            return None
        per = len(cores) // world
        mine = cores[local * per:(local + 1) * per]
        os.sched_setaffinity(0, mine)
        torch.set_num_threads(max(1, min(per, 4)))
        return f'{mine[0]}-{mine[-1]}'
    except Exception:
        return None


def pipeline_tile(B, D):
    """Recurrence tile policy of the pipelined slots: with several batches in flight SM-time counts, not one batch's latency -- 128
    sequences per cluster (the four-sub-tile kernel, half the CTAs per layer launch for 1.2 x the time) where the batch holds whole
    tiles; one batch at a time keeps the automatic 64."""
    return 128 if D > 1 and B % 128 == 0 and os.environ.get('MP_BENCH_TILE', '') != '64' else 0


def measure_cfg3(net, mp, xs, xs_host, B, T, D, args, dist, world, physics):
    """Pipelined device-resident `value` and host-buffer `e2e` of one cfg3 configuration (K8 on or off)."""
    n_sets = len(xs)
    net.enable_physics(physics)
    frames_per_step = B * T * world
    tile = pipeline_tile(B, D)
    pipes = [mp.HostOffline(net, B, T, rec_tile=tile) for _ in range(D)]

    def pipe_step(i):
        pipes[i % D].submit_device(xs[i % n_sets])

    # every slot needs two calls (eager, then graph capture) before it replays its graph
    ms, n = timed_device_steps(pipe_step, args.steps, max(args.warmup, 3 * D), dist, streams=[p.stream for p in pipes],
                               min_seconds=args.min_seconds)
    launches = pipes[0].last_launches
    del pipes
    out = {'value': frames_per_step * n / (ms / 1e3), 'unit': 'frames/s', 'ms_per_step': ms / n, 'steps_timed': n,
           'timed_region_s': ms / 1e3, 'gpu_launches_per_step': launches}

    # e2e: host buffers through the C ABI, copies inside the timed region; depth-D pipeline over batches (HostOffline objects with
    # their own net handle, stream, staging and pinned outputs): batch i+1 is submitted before batch i is awaited
    hosts = [mp.HostOffline(net, B, T, rec_tile=tile) for _ in range(D)]
    slots = {'hosts': hosts}

    def body(n_steps, off, pipelined):
        hosts = slots['hosts']
        if not pipelined:
            for i in range(n_steps):
                hosts[0].run(xs_host[(off + i) % n_sets], None)
            return
        for i in range(n_steps):
            h = hosts[i % D]
            h.wait()                                   # the slot's previous batch is on the host
            h.submit(xs_host[(off + i) % n_sets], None)
        for h in hosts:
            h.wait()

    def e2e_time(pipelined, min_seconds):
        body(max(3, args.warmup, 3 * D if pipelined else 0), 0, pipelined)

        def region(n_steps):
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            t0 = time.perf_counter()
            body(n_steps, 3, pipelined)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if dist is not None:
                t = torch.tensor([dt], device='cuda')
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = t.item()
            return dt

        dt, n_steps = region(args.steps), args.steps
        if min_seconds > 0 and dt < min_seconds:
            reps = int(min(1000, -(-min_seconds // max(dt, 1e-6))))
            if dist is not None:
                t = torch.tensor([reps], device='cuda')
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                reps = int(t.item())
            n_steps = reps * args.steps
            dt = region(n_steps)
        return dt, n_steps

    sync_s, sync_n = e2e_time(False, 0.0)
    e2e_s, e2e_n = e2e_time(True, args.min_seconds)
    out['e2e'] = {'value': frames_per_step * e2e_n / e2e_s, 'unit': 'frames/s', 'h2d_bytes_per_step': B * T * 60 * 4,
                  'd2h_bytes_per_step': B * T * (216 + 72 + 3 + 2) * 4, 'ms_per_step': e2e_s / e2e_n * 1e3, 'steps_timed': e2e_n,
                  'how': f'depth-{D} pipeline over batches through mp_net_enqueue_offline_host (HostOffline.submit / wait): pinned host imu '
                         'in, pose/joints/tran/contact out to pinned host memory, every step',
                  'one_batch_at_a_time': {'value': frames_per_step * sync_n / sync_s, 'unit': 'frames/s', 'ms_per_step': sync_s / sync_n * 1e3}}
    # the same with the compact pose transfer ([B*T,16,6]: 384 instead of 864 B per frame; model_utils.local6d_to_pose rebuilds the
    # matrices on the consumer's side): what a host link shared by 8 GPUs can carry
    del hosts
    slots['hosts'] = None                  # frees the full-transfer slots (3.3 GB of workspace each) before the compact ones exist
    slots['hosts'] = [mp.HostOffline(net, B, T, rec_tile=tile, compact=True) for _ in range(D)]
    c_s, c_n = e2e_time(True, min(args.min_seconds, 1.5))
    out['e2e_compact'] = {'value': frames_per_step * c_n / c_s, 'unit': 'frames/s', 'h2d_bytes_per_step': B * T * 60 * 4,
                          'd2h_bytes_per_step': B * T * (96 + 72 + 3 + 2) * 4, 'ms_per_step': c_s / c_n * 1e3, 'steps_timed': c_n,
                          'how': 'as e2e, pose transferred as the first two columns of the 16 non-ignored joints (HostOffline(compact=True))'}
    slots['hosts'] = None
    return out


def measure_cfg4(net, dev, dist, rank, world, steps, warmup, min_seconds):
    """BASELINE.json config 4: the evaluate.py-shaped synthetic DIP set (50 sequences x 3000 frames) through `evaluate_pose`,
    sequences sharded over the ranks (strong scaling: the set is fixed), every rank's shard in ONE batched forward_offline
    (batch_size = ceil(50 / world)), metric rows on the device, one all-gather of the [50, 8, 2] rows at the end of every pass.
    Host data in (the dataset lives on the host like evaluate.py's), the table out: an end-to-end figure."""
    import hashlib
    from mobileposer_b200.evaluate import evaluate_pose, synthetic_dip
    from mobileposer_b200.sharding import shard_sequences
    items = [(imu.pin_memory(), pose, joint, tran) for imu, pose, joint, tran in synthetic_dip()]    # the dataset lives in pinned host memory
    n = len(items)
    lengths = [it[0].shape[0] for it in items]
    shards = shard_sequences(lengths, world)
    bs = max(len(s) for s in shards)
    physics = net.dynamics_optimizer is not None
    net.enable_physics(False)                         # evaluate.py's default PHYSICS=0: the pinned path

    def one_pass():
        return evaluate_pose(net, items, verbose=False, batch_size=bs)

    for _ in range(max(2, min(warmup, 3))):
        table = one_pass()

    def region(k):
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(k):
            tab = one_pass()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = t.item()
        return dt, tab

    k = max(2, min(steps, 10))
    dt, table = region(k)
    if min_seconds > 0 and dt < min_seconds:
        k = int(min(200, k * -(-min_seconds // max(dt, 1e-6))))
        if dist is not None:
            t = torch.tensor([k], device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            k = int(t.item())
        dt, table = region(k)
    same = True
    if dist is not None:                              # every rank must hold the same gathered table
        ref = table.clone()
        dist.broadcast(ref, 0)
        flag = torch.tensor([float(torch.equal(torch.nan_to_num(ref), torch.nan_to_num(table)))], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        same = bool(flag.item())
    net.enable_physics(physics)
    mean = table.mean(dim=0).cpu()
    rounded = torch.nan_to_num(table, nan=-1.0).cpu().mul(100).round().to(torch.int64).numpy().tobytes()
    frames = sum(lengths)
    return {'workload': workload_name('cfg4', n), 'value': frames * k / dt, 'unit': 'frames/s', 'scaling': 'strong',
            'ms_per_pass': dt / k * 1e3, 'passes_timed': k, 'n_gpus': world, 'sequences': n, 'frames_per_pass': frames,
            'shard_sizes': [len(s) for s in shards], 'batch_size': bs,
            'recurrence_path': ('tcgen05 throughput kernels' if bs * 2 > 30 else
                                'latency cluster kernels: <= 15 sequences per rank never reach the tensor-core tile policy'),
            'all_gather_shape': [n, 8, 2], 'all_gather_block_per_rank': [bs, 16], 'collectives_per_pass': 1 if dist is not None else 0,
            'table_identical_on_all_ranks': same,
            'table_mean': [[round(float(v), 4) for v in row] for row in mean.tolist()],
            'table_sha16_rounded_1e-2': hashlib.sha256(rounded).hexdigest()[:16],
            'how': 'evaluate_pose(model, synthetic_dip(), batch_size=ceil(50 / world)): host IMU in, forward_offline (PHYSICS=0) per '
                   'shard, device metric rows, one all_gather_into_tensor; wall clock with synchronize + barrier, max over ranks'}


def run_ours(args):
    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl ours needs a CUDA device (there is no CPU fallback)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    cores = bind_rank_to_cores(local, world)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist_mod.init_process_group('nccl', device_id=dev)
        dist = dist_mod

    import mobileposer_b200 as mp
    from mobileposer_b200 import _cabi
    from mobileposer_b200.synthetic import synthetic_imu_batch

    T = T_FRAMES
    net, sd = seeded_state_dict()
    net = net.to(dev)
    net.reuse_outputs = True
    peak, peak_src = measured_peak_gbs()

    if args.workload == 'cfg4':
        with ClockSampler(local) as clocks:
            c4 = measure_cfg4(net, dev, dist, rank, world, args.steps, args.warmup, args.min_seconds)
        if rank == 0:
            line = {'metric': METRIC, 'value': c4['value'], 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                    'ms_per_step': c4['ms_per_pass'], 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32',
                    'data': 'synthetic', 'config': bench_config(args, 50, False),
                    'e2e': {'value': c4['value'], 'unit': 'frames/s', 'h2d_bytes_per_step': c4['frames_per_pass'] * 240,
                            'd2h_bytes_per_step': 0, 'note': 'the dataset is host-resident and copied in every pass; only the gathered rows are read back'},
                    'gpu_launches': int(_cabi.lib().mp_launch_count()) * c4['passes_timed'], 'cfg4': c4, 'clocks': clocks.summary(),
                    'all_gather_shape': c4['all_gather_shape']}
            emit(json.dumps(line))
        if dist is not None:
            dist.destroy_process_group()
        return

    B = args.batch or (256 if args.workload == 'cfg3' else 1)
    lens = [T] * B
    phys = physics_on(args)
    net.enable_physics(phys)

    # several distinct input sets so a step never finds its inputs in L2 from the previous step
    n_sets = 8 if B > 1 else 64
    base = rank * B * n_sets
    xs_host = [synthetic_imu_batch(list(range(base + i * B, base + (i + 1) * B)), T).pin_memory() for i in range(n_sets)]
    xs = [x.to(dev) for x in xs_host]

    def step(i):
        net.velocity.rnn_state = None
        return net.forward_offline(xs[i % n_sets], lens)

    # `value`: K steps (then a sustained region, --min-seconds), inputs resident in HBM, D batches in flight (pipeline over batches:
    # D net handles on D streams, so the tail of one batch -- K8, the small GEMMs, the SMs the cluster kernels leave idle --
    # overlaps the head of the next).  The one-batch-at-a-time figure is reported beside it.
    D = args.depth if args.depth > 0 else (6 if args.workload == 'cfg3' else 1)
    frames_per_step = B * T * world
    with ClockSampler(local) as clocks:
        head = measure_cfg3(net, mp, xs, xs_host, B, T, D, args, dist, world, phys)
        ms_seq, n_seq = timed_device_steps(step, args.steps, args.warmup, dist)
        pinned = None
        if phys:      # the reference's default path (PHYSICS=0): the one whose parity is pinned to the reference
            pinned = measure_cfg3(net, mp, xs, xs_host, B, T, D, args, dist, world, False)
            pinned['workload'] = workload_name(args.workload, B, False)
            ms0, n0 = timed_device_steps(step, args.steps, args.warmup, dist)
            pinned['one_batch_at_a_time'] = {'value': frames_per_step * n0 / (ms0 / 1e3), 'unit': 'frames/s', 'ms_per_step': ms0 / n0}
            net.enable_physics(True)
    launches = head['gpu_launches_per_step']
    sequential = {'value': frames_per_step * n_seq / (ms_seq / 1e3), 'unit': 'frames/s', 'ms_per_step': ms_seq / n_seq,
                  'how': 'one batch at a time (MobilePoserNet.forward_offline in a loop, one stream, latency tile policy)'}

    # ---- per-kernel durations (CUDA events on the launching streams), same steps, graphs bypassed ----
    # (through one pipeline slot, i.e. with the tile policy of the headline number; one batch at a time so that a kernel's
    #  duration is its own and not its wait for SMs held by another batch)
    lib = _cabi.lib()
    prof_pipe = mp.HostOffline(net, B, T, rec_tile=pipeline_tile(B, D))
    prof_pipe.submit_device(xs[0])
    prof_pipe.wait()
    _cabi.check(lib.mp_profile_enable(1))
    for i in range(args.steps):
        prof_pipe.submit_device(xs[i % n_sets])
        prof_pipe.wait()
    prof = _cabi.profile_collect()
    _cabi.check(lib.mp_profile_enable(0))
    del prof_pipe
    # dominant kernel: the H=256 recurrence (tcgen05 variant for large batches, FFMA cluster kernel otherwise)
    dom_name = max((k for k in prof if k.startswith('lstm_rec') and 'h64' not in k), key=lambda k: prof[k]['total_ms'], default=None)
    dom = prof.get(dom_name)
    roofline = None
    if dom:
        gbs = dom['algorithmic_bytes'] / (dom['total_ms'] / 1e3) / 1e9
        roofline = {'bound': 'hbm', 'kernel': {'lstm_rec_f16_h256': ('lstm_rec_f16w_kernel (tcgen05 3xFP16 split, W_hh hi+lo in TMEM, 128 sequences per 8-CTA cluster)'
                                                                     if pipeline_tile(B, D) == 128 else
                                                                     'lstm_rec_f16_kernel<N> (tcgen05 3xFP16 split, W_hh hi+lo in TMEM)'),
                                               'lstm_rec_tc_h256': 'lstm_rec_tc_kernel<N> (tcgen05 3xTF32, W_hh in TMEM)',
                                               'lstm_rec_h256': 'lstm_rec_kernel<256,8,*> (FFMA2, W_hh in registers)'}.get(dom_name, dom_name),
                    'achieved': gbs, 'peak': peak, 'unit': 'GB/s', 'frac': gbs / peak, 'traffic': load_traffic(dom_name),
                    'peak_source': peak_src, 'avg_launch_ms': dom['total_ms'] / dom['launches'],
                    'algorithmic_bytes_per_launch': dom['algorithmic_bytes'] / dom['launches'],
                    'launches_per_step': dom['launches'] / args.steps,
                    'note': 'the recurrence is serial in T: bound by the per-step MMA + activation + DSMEM-exchange chain, not by '
                            'HBM (DESIGN.md 4.1); the HBM fraction is the contract figure of the brief'}
        # the launches of this kernel overlap in the pipelined step (6 batches in flight, each launch holds 32 or 16 of the 148 SMs with
        # the 128-sequence tile): what the kernel type moves per unit of wall time is the per-launch figure x the launches in flight
        per_step_ms = dom['total_ms'] / args.steps
        roofline['concurrency'] = {
            'launches_in_flight_avg': per_step_ms / head['ms_per_step'],
            'aggregate_GBps': dom['algorithmic_bytes'] / args.steps / (head['ms_per_step'] / 1e3) / 1e9,
            'aggregate_frac': dom['algorithmic_bytes'] / args.steps / (head['ms_per_step'] / 1e3) / 1e9 / peak,
            'ctas_per_bidirectional_launch': (2 * 8 * (B // 128)) if pipeline_tile(B, D) == 128 else None,
            'note': 'avg_launch_ms is measured one batch at a time through a pipeline slot (same tile policy as the timed region)'}
        if dom_name in ('lstm_rec_tc_h256', 'lstm_rec_f16_h256'):
            f16 = dom_name == 'lstm_rec_f16_h256'
            tf = tensor_peak_tflops() * (2.0 if f16 else 1.0)
            # every launch of this kernel type in the step: sum over layers of 3 products * 2 * B*T*dirs*4H*H flops
            flops = 3 * 2 * sum_recurrent_macs(B, T) * args.steps
            ach = flops / (dom['total_ms'] / 1e3) / 1e12
            roofline['tensor'] = {'achieved_tflops_3_products': ach, 'peak_tflops': tf, 'frac': ach / tf if tf else None,
                                  'peak_source': 'MEASURED_PEAKS.json bf16_tflops' + ('' if f16 else ' / 2 (TF32 runs at half the bf16 rate)')}
    whole = (WEIGHT_BYTES + B * T * IO_BYTES_PER_FRAME) / (head['ms_per_step'] / 1e3) / 1e9 / world
    kernels = {k: {'launches_per_step': v['launches'] / args.steps, 'ms_per_step': v['total_ms'] / args.steps,
                   'algorithmic_GBps': v['algorithmic_bytes'] / (v['total_ms'] / 1e3) / 1e9} for k, v in prof.items()}

    # ---- batch-1 latency configuration (cfg2) on every rank's GPU (replicas) ----------------------------------------
    net.enable_physics(False)      # cfg2 / cfg5 do not name the optimizer
    batch1 = None
    if args.workload == 'cfg3':
        x1 = [synthetic_imu_batch([90000 + rank * 64 + i], T).to(dev) for i in range(16)]

        def step1(i):
            net.velocity.rnn_state = None
            return net.forward_offline(x1[i % 16], [T])
        k1 = max(args.steps, 50)
        ms1, n1 = timed_device_steps(step1, k1, max(args.warmup, 5), dist, min_seconds=min(args.min_seconds, 1.0))
        batch1 = {'workload': workload_name('cfg2', 1), 'value': world * T * n1 / (ms1 / 1e3), 'unit': 'frames/s',
                  'ms_per_step': ms1 / n1, 'steps': n1, 'gpu_launches': net.last_launches,
                  'hbm_roofline_frac_whole_path': (WEIGHT_BYTES + T * IO_BYTES_PER_FRAME) / (ms1 / n1 / 1e3) / 1e9 / peak}

    # ---- live-demo streaming (cfg5): 5 device-combo streams, sliding window, per-tick latency --------------
    streaming = None
    if args.workload == 'cfg3' and rank == 0:
        streaming = {f'window_{W}': streaming_latency(net, dev, W) for W in (45, 125)}
        streaming['note'] = ('per tick: window shift + 4 heads over the whole window + K5 + state update for 5 concurrent '
                             'streams (combos lw_rp, rw_rp, lw_lp, rw_lp, lw_rp_h); window 45 is the reference default '
                             '(config.py:52-54), 125 is BASELINE.json config 5; 60 Hz budget = 16.7 ms')

    # ---- cfg4: sharded evaluate_pose with the all-gather of the [50, 8, 2] rows (the path's only collective) -----------
    cfg4 = None
    if args.workload == 'cfg3' and not args.no_cfg4:
        cfg4 = measure_cfg4(net, dev, dist, rank, world, args.steps, args.warmup, min(args.min_seconds, 2.0))

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = time_cpu_sample(sd, xs_host[0][:min(B, args.cpu_sample)].clone(), physics=phys)
        if phys:
            cpu['pinned_path'] = time_cpu_sample(sd, xs_host[0][:min(B, args.cpu_sample)].clone(), budget_s=10.0, physics=False)
        cpu['small_configs'] = cpu_small_configs(sd)

    if rank == 0:
        line = {
            'metric': METRIC, 'value': head['value'], 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': head['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': bench_config(args, B, phys),
            'setup': {'frames_per_step': frames_per_step, 'steps_timed': head['steps_timed'], 'timed_region_s': head['timed_region_s'],
                      'sustained': f'after the contract region of {args.steps} steps the same steps repeat until the timed region lasts '
                                   f'>= {args.min_seconds} s; value, ms_per_step and the clock median are taken over that region',
                      'l2': f'{n_sets} distinct resident input sets rotate between steps; per-step intermediates '
                            f'({net_workspace_mb(net, B, T):.0f} MB) exceed the 126 MB L2',
                      'parallelism': f'{world} x (one process per GPU, sequences sharded, no data-path collective); '
                                     f'{D} batches in flight per GPU (pipeline over steps)',
                      'recurrence_tile': pipeline_tile(B, D) or 'auto (64 sequences per cluster)',
                      'rank_core_binding': cores},
            'e2e': head['e2e'], 'e2e_compact': head['e2e_compact'], 'gpu_launches': launches * head['steps_timed'], 'gpu_launches_per_step': launches,
            'roofline': roofline, 'whole_path_algorithmic_GBps': whole, 'whole_path_hbm_frac': whole / peak,
            'kernels': kernels, 'one_batch_at_a_time': sequential, 'pinned_path': pinned, 'cpu_baseline': cpu, 'batch1': batch1,
            'streaming': streaming, 'cfg4': cfg4, 'clocks': clocks.summary(),
            'all_gather_shape': cfg4['all_gather_shape'] if (cfg4 and dist is not None) else None,
        }
        emit(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def streaming_latency(net, dev, W, ticks=240, warm=20):
    import mobileposer_b200 as mp
    from mobileposer_b200.synthetic import synthetic_imu_batch
    combos = ['lw_rp', 'rw_rp', 'lw_lp', 'rw_lp', 'lw_rp_h']
    x = torch.stack([synthetic_imu_batch([70000], ticks + warm, combo=c)[0] for c in combos]).to(dev)   # [5, n, 60]
    net.velocity.rnn_state = None
    st = mp.OnlineStreams(net, len(combos), dev, window=W)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(ticks)]
    for t in range(warm):
        st.step(x[:, t])
    torch.cuda.synchronize()
    for t in range(ticks):
        ev[t][0].record()
        st.step(x[:, warm + t])
        ev[t][1].record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    net.velocity.rnn_state = None
    return {'streams': len(combos), 'ticks': ticks, 'p50_ms': ms[len(ms) // 2], 'p99_ms': ms[int(len(ms) * 0.99) - 1],
            'max_ms': ms[-1], 'frames_per_s': len(combos) * 1e3 / (sum(ms) / len(ms))}


def tensor_peak_tflops():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['bf16_tflops']) / 2.0
    except Exception:
        return 1590.0 / 2.0


def sum_recurrent_macs(B, T):
    """W_hh h MACs of the H=256 layers per forward: joints and pose 2 layers x 2 directions, velocity 2 x 1."""
    per_layer_dir = B * T * 4 * 256 * 256
    return per_layer_dir * (4 + 4 + 2)


def net_workspace_mb(net, B, T):
    from mobileposer_b200 import _cabi
    return _cabi.lib().mp_net_workspace_bytes(net._net_handle(), B, T) / 1e6


def load_traffic(kernel):
    """dram bytes per launch from the committed ncu capture summary (profiles/ncu_summary.json), else null."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'ncu_summary.json')) as f:
            return json.load(f)[kernel]['dram_bytes_per_launch']
    except Exception:
        return None


class StdoutGuard:
    """The driver wants ONE JSON line on stdout; NCCL / torchrun children chat there too (e.g. 'NCCL version ...').
    Route everything written to fd 1 to stderr for the duration of the run and keep the real stdout for the line."""

    def __enter__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, line):
        os.write(self.real, (line + '\n').encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.real, 1)
        os.close(self.real)


OUT = None


def emit(line):
    if OUT is not None:
        OUT.emit(line)
    else:
        print(line, flush=True)


if __name__ == '__main__':
    a = parse_args()
    with StdoutGuard() as OUT:
        if a.impl == 'reference':
            run_reference(a)
        else:
            run_ours(a)
