#!/usr/bin/env python
"""bench.py — particle-steps/s of the fp64 LLG Heun ensemble (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): 1,000,000
independent realisations of one 12 nm magnetite particle (K=4e4 J/m^3, Ms=400 kA/m, T=300 K,
alpha=0.1, easy axis and initial moment along z) under a sinusoidal applied field
(f=300 kHz, H0=20 kA/m), explicit Heun, in-kernel Philox noise.  One bench "step" = one pass
of the hot path over that batch: every realisation is advanced by 100,000 Heun steps
(dt=1e-12 s, i.e. 1e-7 s of the field cycle) with 101 zero-order-hold samples and the fused
ensemble reduction — 1e11 particle-steps per pass.  N>1 (one process per GPU): the 1M realisations
are shared out by member index (`--scaling strong`, the default: BASELINE's ensemble is 1M in total;
`--scaling weak` gives every GPU 1M), every member keeps its global Philox index, and ONE
ncclAllReduce — issued by libmagpy_b200 itself on the integration stream, no torch on this arm —
combines the [S][4] ensemble sums per pass.  The other scaling mode is measured too
(config.other_scaling_mode).

value  : whole-job particle-steps/s with inputs resident in HBM, timed with CUDA events on the
         launching stream (max over ranks).
e2e    : same metric through the public API (`EnsembleModel.simulate`) with HOST buffers:
         per pass, seeds/parameters go host->device and final states + ensemble sums come back.
roofline: the integration kernel against the FP64 pipe: W_alg = 98 flop per particle-step
          over the kernel's CUDA-event duration, against the device's FP64
         rate measured in this run: the larger of the library's register-resident DFMA-chain and
         DMMA.8x8x4-chain kernels — one datapath (MEASURED_PEAKS.json holds no fp64 figure).
cpu_baseline / --impl reference: the UNMODIFIED reference `simulation::full_dynamics`
         (compiled into oracle/_ref from its own sources) over a bounded sample of the same
         workload on the host cores, one realisation per call, fanned out by an OpenMP pragma
         that lives in our harness (the reference itself has no threads; its fan-out is joblib).
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

METRIC = 'particle-steps/s (fp64 LLG Heun ensemble)'
UNIT = 'particle-steps/s'
def w_alg(n_particles):
    """fp64 flop per Heun particle-step: 98 + 36 (N - 1)."""
    return 98.0 + 36.0 * (n_particles - 1)


# The default workload (c3) is the configuration the BASELINE metric is quoted on.  `--workload c4` runs BASELINE
# config 4 (64-particle random-geometry clusters, all-pairs dipolar field, Heun) through the same contract — the
# second kernel family (cluster_mma.cu) with its own roofline line; the driver never passes it.
WORKLOADS = {
    'c3': dict(
        name='C3: 1M x 1-particle, sine field 300 kHz / 20 kA/m, Heun dt=1e-12 s, 100k steps per pass',
        N=1, R=1_000_000, radius=12e-9, anisotropy=4e4, Ms=4e5, alpha=0.1, T=300.0,
        dt=1e-12, t_end=1e-7, S=101, field_shape='sine', H0=2e4, f=3e5, random_state=1001),
    'c4': dict(
        name='C4: 100k x 64-particle random clusters (min. separation 30 nm), all-pairs dipolar, Heun dt=1e-14 s, '
             '2000 steps per pass',
        N=64, R=100_000, radius=12e-9, anisotropy=4e4, Ms=4e5, alpha=0.1, T=300.0,
        dt=1e-14, t_end=2e-11, S=21, field_shape='constant', H0=0.0, f=0.0, random_state=1001),
    # the two small BASELINE configurations (latency bound: a thousand members cannot fill a B200) — same contract line
    'c1': dict(
        name='C1: 1000 x 1-particle (12 nm magnetite), no field, Heun dt=1e-14 s to 1e-9 s (1e5 steps), 1000 samples, '
             'per-member trajectories returned',
        N=1, R=1000, radius=12e-9, anisotropy=4e4, Ms=4e5, alpha=0.1, T=300.0, m0=[1.0, 0.0, 0.0],
        dt=1e-14, t_end=1e-9, S=1000, field_shape='constant', H0=0.0, f=0.0, random_state=1001, traj=True),
    'c2': dict(
        name='C2: 10k x 2 dipolar-coupled 7 nm particles 9 nm apart, implicit midpoint (reference quasi-Newton, eps 1e-9), '
             'dt=1e-12 s, 1000 steps, 500 samples',
        N=2, R=10_000, radius=7e-9, anisotropy=1e5, Ms=4e5, alpha=0.1, T=330.0, implicit=True,
        dt=1e-12, t_end=1e-9, S=500, field_shape='constant', H0=0.0, f=0.0, random_state=1001),
}
WORKLOAD = WORKLOADS['c3']


def workload_arrays(R):
    w = WORKLOAD
    N = w['N']
    if N == 1:
        return dict(radius=np.array([w['radius']]), anisotropy=np.array([w['anisotropy']]),
                    axis=np.array([[0.0, 0.0, 1.0]]), m0=np.array([w.get('m0', [0.0, 0.0, 1.0])]), location=np.zeros((1, 3)))
    if N == 2:   # the dimer of docs/source/notebooks/two-particle-equilibrium.ipynb
        z = np.array([[0.0, 0.0, 1.0]] * 2)
        return dict(radius=np.full(2, w['radius']), anisotropy=np.full(2, w['anisotropy']), axis=z, m0=z.copy(),
                    location=np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 9e-9]]))
    from magpy_b200 import geometry   # host-side input generation only (numpy)
    axes = geometry.uniform_random_axes(N, rng=4)
    return dict(radius=np.full(N, w['radius']), anisotropy=np.full(N, w['anisotropy']), axis=axes, m0=axes.copy(),
                location=geometry.random_cluster_coordinates(N, 3e-8, rng=4))


def member_seeds(R, random_state):
    # magpy/model.py:202-203
    np.random.seed(random_state)
    return np.random.randint(np.iinfo(np.int32).max, size=R)


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu')

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None
        self.t_mark = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.device), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '40'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(',')]))

    def mark(self):
        """Start of the timed region: samples from here on are 'timed', earlier ones (the warm-up passes of the same
        workload) are only used when the timed region is too short to catch any."""
        self.t_mark = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

        def digest(rows):
            sm, mx, reasons = [], [], set()
            for r in rows:
                try:
                    sm.append(float(r[1])); mx.append(float(r[2]))
                except (ValueError, IndexError):
                    continue
                for name, val in zip(names, r[5:9]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
            return sm, mx, reasons
        timed = [r for t, r in self.rows if self.t_mark is None or t >= self.t_mark]
        sm, mx, reasons = digest(timed)
        window = 'timed passes'
        if len(sm) < 3:   # a timed region of a few tens of ms: add the samples of the warm-up passes of the same workload
            sm, mx, reasons = digest([r for _, r in self.rows])
            window = 'warm-up + timed passes (the timed region alone caught %d samples)' % len(timed)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons), 'window': window, 'period_ms': 40}


# --------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference on the host cores
# --------------------------------------------------------------------------------------
def load_cpu_reference():
    """(kind, callable) — compiled reference if oracle/_ref exists, else the oracle port."""
    import ctypes as C
    ref = os.path.join(ROOT, 'oracle', '_ref', 'libmagpy_ref.so')
    orc = os.path.join(ROOT, 'oracle', 'libsllg_oracle.so')
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    w = WORKLOAD
    arr = workload_arrays(1)
    field_code = {'sine': 0, 'square': 1, 'constant': 2}[w['field_shape']]
    if os.path.exists(ref):
        lib = C.CDLL(ref)
        lib.ref_ensemble.restype = C.c_double
        # all host cores this process may use, whatever OMP_NUM_THREADS says (torchrun sets it to 1)
        cores = _host_cores()

        def run(seeds):
            seeds = np.ascontiguousarray(seeds, dtype=np.int64)
            sums = np.zeros((w['S'], 4))
            el = lib.ref_ensemble(
                C.c_size_t(len(seeds)), P(seeds), C.c_size_t(w['N']), P(arr['radius']), P(arr['anisotropy']),
                P(arr['axis']), C.c_size_t(0), P(arr['m0']), C.c_size_t(0), P(arr['location']),
                C.c_double(w['Ms']), C.c_double(w['alpha']), C.c_double(w['T']), C.c_int(0), C.c_int(1),
                C.c_int(int(w.get('implicit', False))), C.c_double(1e-9), C.c_double(w['dt']), C.c_double(w['t_end']), C.c_size_t(w['S']),
                C.c_int(field_code), C.c_double(w['H0']), C.c_double(w['f']), C.c_int(cores), P(sums), None)
            if el < 0:
                raise RuntimeError('reference ensemble failed')
            return el
        return 'reference', cores, run
    if not os.path.exists(orc):
        subprocess.check_call(['make', '-C', os.path.join(ROOT, 'oracle')])
    cores = _host_cores()
    os.environ['OMP_NUM_THREADS'] = str(cores)   # read by libgomp when the oracle is loaded
    lib = C.CDLL(orc)
    lib.orc_ensemble.restype = C.c_double

    def run(seeds):
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        sums = np.zeros((w['S'], 4))
        t0 = time.perf_counter()
        lib.orc_ensemble(
            C.c_size_t(len(seeds)), P(seeds), C.c_int(w['N']), P(arr['radius']), P(arr['anisotropy']), P(arr['axis']),
            C.c_size_t(0), P(arr['m0']), C.c_size_t(0), P(arr['location']), C.c_double(w['Ms']),
            C.c_double(w['alpha']), C.c_double(w['T']), C.c_int(0), C.c_int(1), C.c_int(int(w.get('implicit', False))),
            C.c_double(1e-9), C.c_double(w['dt']), C.c_double(w['t_end']), C.c_size_t(w['S']), C.c_int(field_code),
            C.c_double(w['H0']), C.c_double(w['f']), P(sums), None)
        return time.perf_counter() - t0
    return 'port', cores, run


def _joblib_member(seed, w, arr):
    """One ensemble member the way the reference runs it: one `simulation::full_dynamics` call (magpy/core.pyx:149-185)
    in a worker process, the result dict pickled back to the parent (magpy/model.py:204-207)."""
    import ctypes as C
    lib = _joblib_member.lib = getattr(_joblib_member, 'lib', None) or C.CDLL(os.path.join(ROOT, 'oracle', '_ref', 'libmagpy_ref.so'))
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    N, S = w['N'], w['S']
    t, fl, m = np.zeros(S), np.zeros(S), np.zeros((N, 3, S))
    rc = lib.ref_simulate(C.c_size_t(N), P(arr['radius']), P(arr['anisotropy']), P(arr['axis']), P(arr['m0']), P(arr['location']),
                          C.c_double(w['Ms']), C.c_double(w['alpha']), C.c_double(w['T']), C.c_int(0), C.c_int(1),
                          C.c_int(int(w.get('implicit', False))), C.c_double(1e-9), C.c_double(w['dt']), C.c_double(w['t_end']), C.c_size_t(S), C.c_long(int(seed)),
                          C.c_int({'sine': 0, 'square': 1, 'constant': 2}[w['field_shape']]), C.c_double(w['H0']),
                          C.c_double(w['f']), P(t), P(fl), P(m))
    if rc != 0:
        raise RuntimeError('reference full_dynamics failed')
    return {'N': N, 'time': t, 'field': fl, 'x': {i: m[i, 0] for i in range(N)}, 'y': {i: m[i, 1] for i in range(N)},
            'z': {i: m[i, 2] for i in range(N)}}


def cpu_sample_joblib(cores, n_steps, target_seconds):
    """BASELINE.md section 3 baseline (i), the reference's NATIVE fan-out: a joblib (loky) process pool with one
    `full_dynamics` call per member and the per-member results pickled back, warm pool (first batch discarded), timed
    from the parent.  The member function is the compiled, unmodified reference; the reference's own Python package
    cannot travel to the GPU box, so the pool is driven from here (scripts/reference_native_baseline.py runs the real
    `magpy.EnsembleModel.simulate(n_jobs=...)` in the build container for comparison)."""
    from joblib import Parallel, delayed
    w, arr = WORKLOAD, workload_arrays(1)
    seeds = member_seeds(1 << 16, w['random_state'])
    with Parallel(n_jobs=cores) as pool:
        pool(delayed(_joblib_member)(s, w, arr) for s in seeds[:cores])          # spawns and warms the workers (discarded)
        t0 = time.perf_counter()
        pool(delayed(_joblib_member)(s, w, arr) for s in seeds[:2 * cores])      # calibration
        per_member = (time.perf_counter() - t0) / 2
        n = int(max(cores, min(len(seeds), cores * max(1, round(target_seconds / max(per_member, 1e-3))))))
        t0 = time.perf_counter()
        res = pool(delayed(_joblib_member)(s, w, arr) for s in seeds[:n])
        el = time.perf_counter() - t0
    mz = np.mean([r['z'][0][-1] for r in res])
    return w['N'] * n * n_steps / el, n, el, float(mz)


def _host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def steps_per_member():
    """Steps the reference executes per member for the workload: the schedule of
    lib/simulation.cpp:342-355 evaluated in plain Python (no product code on this path)."""
    w = WORKLOAD
    KB, MU0, GYROMAG = 1.38064852e-23, 1.25663706e-6, 1.76086e11   # include/constants.hpp:10-12
    H_k = 2 * w['anisotropy'] / MU0 / w['Ms']
    tau = GYROMAG * MU0 * H_k / (1 + w['alpha'] * w['alpha'])
    dt, T = w['dt'] * tau, w['t_end'] * tau
    Ts = T / (w['S'] - 1)
    lim = (w['S'] - 1) * Ts
    s = max(int(lim / dt) - 2, 0)
    while s * dt <= lim:
        s += 1
    return s


def cpu_sample(run, cores, n_steps, target_seconds):
    """Time the CPU path on a bounded sample: calibrate, then one batch of ~target_seconds."""
    seeds = member_seeds(1 << 16, WORKLOAD['random_state'])
    n0 = max(cores, 8)
    el0 = run(seeds[:n0])
    rate0 = n0 * n_steps / el0
    n1 = int(min(len(seeds), max(n0, rate0 * target_seconds / n_steps)))
    n1 = max(cores, (n1 // cores) * cores)
    el1 = run(seeds[:n1])
    return WORKLOAD['N'] * n1 * n_steps / el1, n1, el1


def reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    kind, cores, run = load_cpu_reference()
    n_steps = steps_per_member()
    seeds = member_seeds(1 << 16, WORKLOAD['random_state'])
    # bounded sample per step: ~5 s of host work, calibrated once
    el = run(seeds[:max(cores, 8)])
    rate = max(cores, 8) * n_steps / el
    n = int(max(cores, min(len(seeds), rate * 5.0 / n_steps)))
    n = max(cores, (n // cores) * cores)
    for _ in range(args.warmup):
        run(seeds[:n])
    t = 0.0
    for _ in range(args.steps):
        t += run(seeds[:n])
    value = WORKLOAD['N'] * args.steps * n * n_steps / t
    sample = '%d realisations x %d %s steps per step (of the %d-realisation workload), %d host threads' % (
        n, n_steps, 'implicit-midpoint' if WORKLOAD.get('implicit') else 'Heun', WORKLOAD['R'], cores)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * t / args.steps, 'higher_is_better': True,
        'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD['name'], 'sample': sample},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def _make_plan(core, arr, seeds, R_local, offset, device, comm, gauss='f32p', t_end=None):
    w = WORKLOAD
    return core.EnsemblePlan(arr['radius'], arr['anisotropy'], arr['axis'], arr['m0'], arr['location'], w['Ms'],
                             w['alpha'], w['T'], False, True, bool(w.get('implicit', False)), w['dt'], t_end or w['t_end'], w['S'], seeds,
                             field_shape=w['field_shape'], field_amplitude=w['H0'], field_frequency=w['f'],
                             device=device, stream_offset=offset, return_trajectories=bool(w.get('traj', False)),
                             return_sums=True, return_final=True, gauss=gauss, comm=comm)


def _timed_passes(plan, comm, passes):
    """`passes` passes of the hot path; (device ms incl. the all-reduce, integrate-kernel ms, last stats), each the
    max over ranks of the per-rank CUDA-event sums (events on the library's launch stream)."""
    t_dev = t_int = 0.0
    st = None
    for _ in range(passes):
        plan.run()          # integration kernels + ensemble reduction + ONE ncclAllReduce, all on the plan's stream
        st = plan.sync()
        t_dev += st['device_ms']
        t_int += st['integrate_ms']
    tt = np.array([t_dev, t_int])
    if comm is not None:
        comm.allreduce(tt, 'max')
    return float(tt[0]), float(tt[1]), st


def ours(args):
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    import magpy_b200 as mp
    from magpy_b200 import core
    from magpy_b200.sharding import shard_bounds

    if core.device_count() == 0:
        raise RuntimeError('bench.py needs a CUDA device: magpy_b200 has no CPU fallback')
    # one process per GPU; the library's own NCCL communicator (no torch anywhere on this arm)
    comm = core.Comm.from_env(local_rank) if world > 1 else None
    w = WORKLOAD
    strong = args.scaling == 'strong'
    R_total = w['R'] if strong else w['R'] * world
    seeds_all = member_seeds(R_total, w['random_state'])
    lo, hi = shard_bounds(R_total, world, rank)
    R = hi - lo
    arr = workload_arrays(R)

    def barrier():
        if comm is not None:
            comm.barrier()

    # roofline denominator: the best FP64 rate measured on this device in this run.  DFMA and DMMA are the same
    # datapath (scripts/micro/dmma.cu); the DMMA chain reaches the nominal rate, the DFMA chain ~8 % less.
    dfma_tflops, max_mhz = core.fp64_peak(local_rank)
    dmma_tflops = core.fp64_mma_peak(local_rank)
    peak_tflops = max(dfma_tflops, dmma_tflops)

    plan = _make_plan(core, arr, seeds_all[lo:hi], R, lo, local_rank, comm)
    sampler = ClockSampler(local_rank)
    sampler.start()
    _timed_passes(plan, comm, max(args.warmup, 0))
    launches0 = plan.sync()['kernel_launches']
    barrier()
    sampler.mark()
    wall0 = time.perf_counter()
    t_dev, t_int, st = _timed_passes(plan, comm, args.steps)
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    gpu_launches = st['kernel_launches'] - launches0
    n_steps = st['steps_per_member']
    ps_per_pass = R_total * w['N'] * n_steps          # whole job, all ranks
    ms_per_step = t_dev / args.steps
    int_ms = t_int / args.steps
    value = ps_per_pass / (ms_per_step * 1e-3)
    out = plan.fetch()
    mean_mz = float(out['sums'][-1, 2] / R_total / w['Ms'] / w['N'])   # the plan's sums were all-reduced in place
    variant = st.get('kernel_variant', 0)
    del plan

    # the other scaling mode, for the record (N > 1 only; at N = 1 they coincide)
    other = None
    if world > 1:
        R2 = w['R'] if strong else -(-w['R'] // world)
        off2 = rank * R2
        seeds2 = member_seeds(R2 * world, w['random_state'])[off2:off2 + R2]
        plan2 = _make_plan(core, workload_arrays(R2), seeds2, R2, off2, local_rank, comm)
        _timed_passes(plan2, comm, 1)
        barrier()
        k2 = max(1, min(args.steps, 3))
        t2, _, st2 = _timed_passes(plan2, comm, k2)
        other = {'scaling': 'weak' if strong else 'strong', 'realisations_per_gpu': R2, 'ms_per_step': t2 / k2,
                 'value': R2 * world * w['N'] * st2['steps_per_member'] / (t2 / k2 * 1e-3), 'unit': UNIT}
        del plan2

    # the other Gaussian transforms of the Philox stream on the same workload (N = 1 run of the default workload only)
    rng_rates = None
    if world == 1 and w is WORKLOADS['c3']:
        rng_rates = {}
        for g, frac in (('f32', 0.2), ('f64', 0.05)):
            p3 = _make_plan(core, arr, seeds_all[lo:hi], R, lo, local_rank, None, gauss=g, t_end=w['t_end'] * frac)
            p3.run(); p3.sync(); p3.run()
            s3 = p3.sync()
            rng_rates[g] = s3['particle_steps'] / (s3['integrate_ms'] * 1e-3)
            del p3

    # end to end through the public API with host buffers (H2D + D2H inside the timed region)
    base = mp.Model(arr['radius'], arr['anisotropy'], arr['axis'], arr['m0'], arr['location'], w['Ms'], w['alpha'],
                    w['T'], field_shape=w['field_shape'], field_frequency=w['f'], field_amplitude=w['H0'])
    ens = mp.EnsembleModel(R_total, base)
    shard = (rank, world) if world > 1 else None
    e2e_passes = max(1, min(args.steps, 2))
    implicit, traj = bool(w.get('implicit', False)), bool(w.get('traj', False))
    for _ in range(2):   # warm-up: two calls, because the caller holds one result while the next is produced — the pool of
        #            page-locked output blocks reaches its steady state (two blocks per array) with the second call
        res = ens.simulate(w['t_end'], w['dt'], w['S'], w['random_state'], implicit_solve=implicit, device=local_rank,
                           return_trajectories=traj, shard=shard, comm=comm)
    barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    pass_s = []                                                   # wall clock of every pass, for the record (this rank)
    for i in range(e2e_passes):
        tp = time.perf_counter()
        res = ens.simulate(w['t_end'], w['dt'], w['S'], w['random_state'], implicit_solve=implicit,
                           device=local_rank, return_trajectories=traj, shard=shard, comm=comm)
        h2d += sum(s['h2d_bytes'] for s in res.stats)
        d2h += sum(s['d2h_bytes'] for s in res.stats)
        final_mz = float(res.ensemble_magnetisation()[-1])        # the step's result, read on the host
        pass_s.append(time.perf_counter() - tp)
    barrier()
    tt = np.array([(time.perf_counter() - t0) / e2e_passes, float(h2d), float(d2h)])
    if comm is not None:
        e2e_t = comm.allreduce(tt[:1].copy(), 'max')
        tot = comm.allreduce(tt[1:].copy(), 'sum')
        tt = np.concatenate([e2e_t, tot])
    e2e_s, h2d, d2h = float(tt[0]), int(tt[1]), int(tt[2])
    e2e_value = ps_per_pass / e2e_s

    if rank != 0:
        return

    # implicit midpoint has no fixed algorithmic work (the iteration count is data dependent): the
    # line then carries the quasi-Newton iterations per step instead of a roofline fraction
    W_ALG = None if w.get('implicit') else w_alg(w['N'])
    # the dominant kernel of a rank processes that rank's share of the pass
    achieved = W_ALG * (ps_per_pass / world) / (int_ms * 1e-3) / 1e12 if W_ALG else None
    traffic = None
    prof = os.path.join(ROOT, 'profiles', 'heun_single_traffic.json')
    if args.workload == 'c3' and world == 1 and os.path.exists(prof):   # the capture is of the C3 kernel at the C3 size
        try:
            traffic = json.load(open(prof)).get('dram_bytes_per_launch')
        except Exception:
            traffic = None
    cpu = None
    try:
        if world > 1:
            raise RuntimeError('timed at N=1 only (see the N=1 line and --impl reference)')
        kind, cores, run = load_cpu_reference()
        v, n, el = cpu_sample(run, cores, n_steps + 1, 10.0)
        cpu = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': kind,
               'sample': '%d realisations x %d Heun steps of the same workload in %.1f s' % (n, n_steps + 1, el),
               'fanout': 'OpenMP pragma in our harness around the unmodified full_dynamics (BASELINE.md section 3, baseline ii)'}
        if kind == 'reference':
            try:   # baseline (i): the reference's native joblib process pool; `value` stays the larger of the two
                vj, nj, elj, _ = cpu_sample_joblib(cores, n_steps + 1, 8.0)
                cpu['joblib'] = {'value': vj, 'unit': UNIT, 'cores': cores,
                                 'sample': '%d realisations x %d Heun steps in %.1f s, warm loky pool, results pickled '
                                           'back per member (magpy/model.py:204-207)' % (nj, n_steps + 1, elj)}
                if vj > cpu['value']:
                    cpu['value'], cpu['fanout'] = vj, 'joblib process pool (baseline i)'
            except Exception as exc:
                cpu['joblib'] = {'value': None, 'error': str(exc)}
    except Exception as exc:   # the baseline is reported, never required for the GPU number
        cpu = {'value': None, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'unavailable', 'sample': str(exc)}

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': w['name'], 'realisations_total': R_total, 'realisations_per_gpu': -(-R_total // world),
                   'particles': w['N'], 'heun_steps_per_pass': n_steps, 'samples': w['S'],
                   'rng': 'Philox4x32-10 + fp32 Box-Muller in kernel, one Philox block per two steps (gauss=f32p)',
                   'rng_other_modes_particle_steps_per_s': rng_rates,
                   'collective': 'one ncclAllReduce(sum, fp64) of the [S][4] ensemble sums per pass, issued by '
                                 'libmagpy_b200 on the integration stream' if world > 1 else 'none (one GPU)',
                   'l2': 'state is register resident; no input is re-read between passes (0 B/step steady-state HBM '
                         'traffic), so no L2 flush applies',
                   'timing': 'CUDA events on the launching stream (integration + reduction + all-reduce), max over ranks',
                   'kernel_variant': ({1: 'heun_single, free register allocation (6 CTAs/SM)', 7: 'heun_single, 7 CTAs/SM',
                                       100: 'heun_single, latency variant (field table prefetched)',
                                       200: 'heun_single_balanced: persistent kernel over (time segment, 128-member block) tasks',
                                       300: 'heun_single_split: integrator warp + three generator warps per 32 members '
                                            '(at most 64 members per SM)'
                                       }.get(variant, str(variant)) if w['N'] == 1 and not w.get('implicit') else None),
                   'other_scaling_mode': other,
                   'mean_mz_over_Ms_at_end': mean_mz, 'wall_s_timed_region': wall},
        'roofline': {'bound': 'fp64', 'achieved': achieved, 'peak': peak_tflops, 'unit': 'TFLOP/s',
                     'frac': achieved / peak_tflops if (peak_tflops and achieved) else None, 'traffic': traffic,
                     'newton_iterations_per_step': (st['newton_iterations'] / max(1, st['particle_steps'] / w['N'])
                                                    if w.get('implicit') else None),
                     'kernel': ('heun_single_balanced' if variant == 200 else 'heun_single_split' if variant == 300
                                else st['kernel']) + '_kernel',
                     'what_binds': ('instruction issue, not the FP64 lanes: on sm_100a an FP64 instruction holds its sub-partition\'s '
                                    'issue port 2 cycles, 3 with three distinct register operands (scripts/micro/dfma_operands.cu, '
                                    'profiles/r02_dfma_operands.txt); a Heun step pair is 74 FP64 (30 three-register) + 91 other '
                                    'instructions = 269 issue cycles per warp, 290 measured (DESIGN.md section 4); `frac` credits '
                                    'W_alg = 98 flop per step, the kernel executes 37 FP64 instructions per step'
                                    if w['N'] == 1 and not w.get('implicit') and variant == 200 else None),
                     'kernel_ms_per_launch': int_ms,
                     'algorithmic_flop_per_particle_step': W_ALG,
                     'peak_dfma_chain': dfma_tflops, 'peak_dmma_chain': dmma_tflops, 'sm_clock_mhz_max': max_mhz,
                     'peak_source': 'measured in this run: the larger of the library\'s register-resident DFMA-chain and '
                                    'DMMA.8x8x4-chain kernels (one FP64 datapath); MEASURED_PEAKS.json has no fp64 figure',
                     'peak_recompute': 'fp64_peak_kernel: SMs x 8 CTAs x 128 threads, 8192 iterations of 64 independent '
                                       'DFMA per thread -> flop = threads x 8192 x 64 x 2; fp64_mma_peak_kernel: SMs x 256 '
                                       'threads, 4000 iterations of 48 DMMA.8x8x4 per warp -> flop = warps x 4000 x 48 x 512; '
                                       'nominal = SMs x 4 x 16 lanes x 2 x clock'},
        'cpu_baseline': cpu,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d // e2e_passes,
                'd2h_bytes_per_step': d2h // e2e_passes, 'api': 'EnsembleModel.simulate', 'passes': e2e_passes,
                'pass_s': [round(x, 6) for x in pass_s],
                'mean_mz_over_Ms_at_end': final_mz / w['Ms'] / w['N']},
        'gpu_launches': int(gpu_launches),
        'clocks': clocks,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c3', choices=sorted(WORKLOADS),
                    help='c3 (default): the configuration the BASELINE metric is quoted on; c4: 64-particle clusters; '
                         'c1, c2: the two small (latency-bound) BASELINE configurations')
    ap.add_argument('--scaling', default='strong', choices=['strong', 'weak'],
                    help='strong (default): the workload\'s realisations are shared out over the GPUs (BASELINE: 1M in '
                         'total); weak: every GPU takes the full count.  The other mode is measured too and reported in '
                         'config.other_scaling_mode')
    args = ap.parse_args()
    global WORKLOAD
    WORKLOAD = WORKLOADS[args.workload]
    if args.impl == 'reference':
        reference_arm(args)
    else:
        ours(args)


if __name__ == '__main__':
    main()
