"""Host-side logic and the C-ABI surface, runnable without a GPU: exported symbols, the
reference's host arithmetic (unit conversion, schedule) against the oracle, error behaviour,
the Results/EnsembleResults/EnsembleModel mirror of the reference API, and the world_size-2
sharding path over gloo (the device call replaced by the oracle, as a stand-in checker)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as ol

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(ROOT, 'magpy_b200', 'libmagpy_b200.so')


@pytest.fixture(scope='module')
def orc():
    return ol.load_oracle()


@pytest.fixture(scope='module')
def core():
    import magpy_b200.core as core
    return core


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'magpy_b200.h')).read()
    names = sorted(set(re.findall(r'\b(magpy_b200_[a-z0-9_]+)\s*\(', hdr)))
    assert len(names) >= 18
    lib = C.CDLL(LIB)
    for n in names:
        assert hasattr(lib, n), n
    lib.magpy_b200_abi_version.restype = C.c_int
    assert lib.magpy_b200_abi_version() == int(re.search(r'#define MAGPY_B200_ABI_VERSION (\d+)', hdr).group(1))
    # every declared entry cites the reference interface it replaces
    assert hdr.count('lib/') + hdr.count('magpy/') + hdr.count('include/') >= 15


def test_constants(core):
    # include/constants.hpp:10-12 (MU0 is the reference's truncated value, not 4 pi 1e-7)
    assert core.get_KB() == 1.38064852e-23
    assert core.get_mu0() == 1.25663706e-6
    assert core.get_gamma() == 1.76086e11


def test_reduce_units_matches_oracle_bitwise(orc, core):
    rng = np.random.default_rng(5)
    for N in (1, 2, 7, 64):
        c = ol.make_case(N=N, radius=7e-9 * (1 + rng.random(N)), anisotropy=1e5 * (0.5 + rng.random(N)), Ms=4.3e5,
                         alpha=0.07, T=311.0, dt=3e-13, t_end=2e-9, H0=1.7e4, f=2.5e5, rng=rng)
        want = ol.reduced_scalars(orc, c)
        got = core.reduce_units(c.radius, c.anisotropy, c.Ms, c.alpha, c.T, c.dt, c.t_end, c.H0, c.f)
        for k in ('V_av', 'K_av', 'H_k', 'time_factor', 'dt_red', 'T_red', 'h0', 'f_red', 'dipolar_prefactor'):
            assert got[k] == want[k], k
        for k in ('k_red', 'v_red', 'sigma'):
            assert np.array_equal(got[k], want[k]), k


def test_config1_reduced_values(core):
    r = core.reduce_units([12e-9], [4e4], 4e5, 0.1, 300.0, 1e-14, 1e-9)
    assert np.isclose(r['V_av'], 7.238e-24, rtol=1e-3) and np.isclose(r['H_k'], 159154.9, rtol=1e-6)
    assert np.isclose(r['time_factor'], 3.4868e10, rtol=1e-4) and np.isclose(r['sigma'][0], 0.03764, rtol=1e-3)


def test_schedule_matches_literal_loop(orc, core):
    rng = np.random.default_rng(2)
    cases = [(3.4868514851485146e-4, 34.868514851485145, 1000), (0.3, 1.0, 11), (1.0, 10.0, 11), (0.1, 1.0, 11),
             (0.7, 0.5, 5), (1e-3, 1.0, 2), (0.25, 100.0, 401)]
    cases += [(float(rng.uniform(1e-4, 0.5)), float(rng.uniform(0.5, 40)), int(rng.integers(2, 300))) for _ in range(40)]
    for dt, T, S in cases:
        assert np.array_equal(core.schedule(dt, T, S), ol.schedule(orc, dt, T, S)), (dt, T, S)


def test_argument_errors_do_not_need_a_device(core):
    c = ol.make_case(N=2)
    args = (c.radius, c.anisotropy, c.axis, c.m0, c.location, c.Ms, c.alpha, c.T, False, True, False, c.dt, c.t_end)
    with pytest.raises(KeyError):            # magpy/core.pyx:158-161
        core.simulate(*args, 10, 1, field_shape='sawtooth')
    with pytest.raises(ValueError):
        core.simulate(*args, 1, 1)
    with pytest.raises(ValueError):
        core.simulate(c.radius, c.anisotropy[:1], c.axis, c.m0, c.location, c.Ms, c.alpha, c.T, False, True, False,
                      c.dt, c.t_end, 10, 1)
    with pytest.raises(ValueError):
        core.simulate_ensemble(c.radius, c.anisotropy, c.axis[None][:, :1], c.m0, c.location, c.Ms, c.alpha, c.T, False,
                               True, False, c.dt, c.t_end, 10, [1, 2, 3])
    with pytest.raises(KeyError):
        core.simulate_ensemble(c.radius, c.anisotropy, c.axis, c.m0, c.location, c.Ms, c.alpha, c.T, False,
                               True, False, c.dt, c.t_end, 10, [1, 2, 3], gauss='f16')
    with pytest.raises(ValueError):          # single-process multi-GPU entry: an empty device list
        core.simulate_ensemble(c.radius, c.anisotropy, c.axis, c.m0, c.location, c.Ms, c.alpha, c.T, False,
                               True, False, c.dt, c.t_end, 10, [1, 2, 3], devices=[])
    with pytest.raises(ValueError):
        core.simulate_ensemble(c.radius, c.anisotropy, c.axis, c.m0, c.location, c.Ms, c.alpha, c.T, False,
                               True, False, c.dt, c.t_end, 10, [1, 2, 3], devices='some')
    # discrete-orientation model (magpy/core.pyx:205-278): bad shape -> KeyError, too few samples / bad scalars -> ValueError
    p0 = np.array([1.0, 0.0])
    with pytest.raises(KeyError):
        core.simulate_dom(p0, 1e-24, 4e4, 300.0, 4e5, 0.1, 1e-10, 1e-6, 10, 'sawtooth', 0.0, 0.0, 1)
    with pytest.raises(ValueError):
        core.simulate_dom(p0, 1e-24, 4e4, 300.0, 4e5, 0.1, 1e-10, 1e-6, 1, 'constant', 0.0, 0.0, 1)
    with pytest.raises(ValueError):
        core.simulate_dom(p0, 1e-24, 4e4, 300.0, 4e5, 0.1, -1e-10, 1e-6, 10, 'constant', 0.0, 0.0, 1)
    with pytest.raises(ValueError):
        core.simulate_dom_batch(np.ones((3, 2)), np.ones(2), np.ones(3), 300.0, 4e5, 0.1, 1e-10, 1e-6, 10)


def test_no_cpu_fallback(core):
    if core.device_count() > 0:
        pytest.skip('a CUDA device is present')
    c = ol.make_case(N=1)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        core.simulate(c.radius, c.anisotropy, c.axis, c.m0, c.location, c.Ms, c.alpha, c.T, False, True, False, c.dt,
                      c.t_end, 10, 1)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        core.simulate_dom(np.array([1.0, 0.0]), 1e-24, 4e4, 300.0, 4e5, 0.1, 1e-10, 1e-6, 10, 'constant', 0.0, 0.0, 1)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        core.fp64_peak()
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        core.simulate_ensemble(c.radius, c.anisotropy, c.axis, c.m0, c.location, c.Ms, c.alpha, c.T, False, True,
                               False, c.dt, c.t_end, 10, [1, 2, 3], devices=[0, 1])


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, 'magpy_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.pyx', '.cu', '.cuh', '.h', '.cpp')):
                src = open(os.path.join(dirpath, f)).read()
                assert 'sllg_oracle' not in src and 'libmagpy_ref' not in src and 'oracle_lib' not in src, f


def _fake_results(R=6, N=2, S=9, seed=0):
    rng = np.random.default_rng(seed)
    traj = rng.normal(size=(R, N, 3, S))
    time = np.linspace(0, 1e-6, S)
    field = 1e4 * np.sin(2 * np.pi * 2e6 * time)
    M = traj.sum(axis=1)
    sums = np.stack([M[:, 0].sum(0), M[:, 1].sum(0), M[:, 2].sum(0), (M[:, 2] ** 2).sum(0)], axis=1)
    return time, field, traj, sums


def test_results_api_matches_reference_semantics():
    from magpy_b200 import Results, EnsembleResults, get_mu0
    time, field, traj, sums = _fake_results()
    R, N = traj.shape[:2]
    members = [Results(time, field, {p: traj[i, p, 0] for p in range(N)}, {p: traj[i, p, 1] for p in range(N)},
                       {p: traj[i, p, 2] for p in range(N)}, N) for i in range(R)]
    a = EnsembleResults(members)                                   # the reference's constructor
    b = EnsembleResults.from_arrays(time, field, R, trajectories=traj, sums=sums, final=traj[..., -1])
    c = EnsembleResults.from_arrays(time, field, R, sums=sums, final=traj[..., -1])   # 1M-member style: no trajectories
    # magpy/results.py:61-75 (sum over particles), :134-151 (mean over members)
    assert np.allclose(members[0].magnetisation('x'), traj[0, :, 0].sum(0))
    for d in 'xyz':
        want = np.mean([m.magnetisation(d) for m in members], axis=0)
        for e in (a, b, c):
            assert np.allclose(e.ensemble_magnetisation(d), want)
    assert np.allclose(np.array(a.magnetisation()), np.array(b.magnetisation()))
    assert a.final_state()[2]['y'][1] == traj[2, 1, 1, -1] == b.final_state()[2]['y'][1] == c.final_state()[2]['y'][1]
    # magpy/results.py:167-217 with the missing get_mu0 import repaired
    want = -get_mu0() * np.trapezoid(field, b.ensemble_magnetisation())
    for e in (a, b, c):
        assert np.isclose(e.energy_dissipated(), want)
    mask = time >= time[-1] - 1 / 2e6
    want = -get_mu0() * np.trapezoid(field[mask], b.ensemble_magnetisation()[mask])
    assert np.isclose(b.final_cycle_energy_dissipated(2e6), want)
    assert len(b.results) == R and np.array_equal(b.results[-1].z[1], traj[-1, 1, 2])
    with pytest.raises(ValueError):
        c.magnetisation()
    se = b.ensemble_magnetisation_stderr()
    assert np.allclose(se, traj[:, :, 2].sum(1).std(axis=0, ddof=1) / np.sqrt(R))


def test_results_persistence_round_trip(tmp_path):
    import pickle
    from magpy_b200.results import save_results, load_results, EnsembleResults
    time, field, traj, sums = _fake_results(R=4, N=2, S=7)
    ens = EnsembleResults.from_arrays(time, field, 4, trajectories=traj, sums=sums, final=traj[..., -1])
    prefix = str(tmp_path / 'member3')
    save_results(prefix, ens.results[3], particle=1)
    # raw native float64, no header: the reference's io::write_array (include/io.hpp:12-27)
    assert os.path.getsize(prefix + '.mz') == 7 * 8
    back = load_results(prefix)
    assert np.array_equal(back.z[0], traj[3, 1, 2]) and np.array_equal(back.time, time)
    assert np.array_equal(back.field, field) and back.N == 1
    # magpy/data.py shelves pickles of the results: the array-backed ensemble survives a round trip
    again = pickle.loads(pickle.dumps(ens))
    assert np.array_equal(again.ensemble_magnetisation(), ens.ensemble_magnetisation())
    assert np.array_equal(again.results[2].x[1], traj[2, 1, 0])


def test_sar_is_cycle_energy_times_frequency_over_density():
    from magpy_b200.results import EnsembleResults
    from magpy_b200.core import get_mu0
    f, H0, Ms, S = 3e5, 2e4, 4e5, 4001
    t = np.linspace(0, 2 / f, S)
    H = H0 * np.sin(2 * np.pi * f * t)
    M = 0.8 * Ms * np.sin(2 * np.pi * f * t - 0.3)            # lagging response: elliptical loop
    sums = np.zeros((S, 4)); sums[:, 2] = 10 * M
    res = EnsembleResults.from_arrays(t, H, 10, sums=sums)
    area = np.pi * H0 * 0.8 * Ms * np.sin(0.3) * get_mu0()   # mu0 * pi * H0 * M0 * sin(phase lag) per cycle
    # the reference's sign convention (magpy/results.py:191): -mu0 * trapz(field, M) is negative for a lagging M
    assert abs(res.final_cycle_energy_dissipated(f) / -area - 1) < 1e-3
    assert abs(res.specific_absorption_rate(f) / (area * f / 5180.0) - 1) < 1e-3


def test_ensemble_model_seeds_and_grouping(monkeypatch):
    import magpy_b200 as mp
    from magpy_b200 import model as model_mod
    base = mp.Model([7e-9, 7e-9], [1e5, 1e5], [[0, 0, 1.0]] * 2, [[0, 0, 1.0]] * 2, [[0, 0, 0], [0, 0, 9e-9]], 4e5, 0.1, 330.0)
    R = 7
    rng = np.random.default_rng(0)
    axes = rng.normal(size=(R, 2, 3))
    temps = [300.0, 300.0, 310.0, 300.0, 310.0, 320.0, 300.0]
    ens = mp.EnsembleModel(R, base, anisotropy_axis=list(axes), temperature=temps)
    assert len(ens.models) == R and ens.models[2].temperature == 310.0      # magpy/model.py:146-156
    calls = []

    def fake(radius, anisotropy, axis, m0, location, Ms, alpha, T, renorm, inter, impl, dt, t_end, S, seeds, shape,
             H0, f, tol, **kw):
        calls.append(dict(T=T, seeds=np.array(seeds), axis=np.array(axis), kw=kw))
        n = len(seeds)
        return {'N': 2, 'R': n, 'time': np.arange(S) * 1.0, 'field': np.zeros(S),
                'trajectories': np.ones((n, 2, 3, S)) * T, 'sums': np.ones((S, 4)) * n, 'final': np.ones((n, 2, 3)) * T,
                'stats': {}}
    monkeypatch.setattr(model_mod.core, 'simulate_ensemble', fake)

    class FakePlan:   # several parameter groups run as concurrent plans: every plan is enqueued before the first sync
        events = []

        def __init__(self, *args, **kw):
            self.out = fake(*args, **kw)

        def run(self):
            FakePlan.events.append('run')

        def sync(self):
            FakePlan.events.append('sync')
            return {'kernel': 'fake'}

        def fetch(self):
            return dict(self.out)
    monkeypatch.setattr(model_mod.core, 'EnsemblePlan', FakePlan)
    res = ens.simulate(1e-9, 1e-12, 5, random_state=42, implicit_solve=False, n_jobs=8)
    assert FakePlan.events == ['run'] * 3 + ['sync'] * 3
    # seeds exactly as magpy/model.py:202-203
    np.random.seed(42)
    want = np.random.randint(np.iinfo(np.int32).max, size=R)
    assert sorted(c['T'] for c in calls) == [300.0, 310.0, 320.0]
    for c in calls:
        idx = [i for i, t in enumerate(temps) if t == c['T']]
        assert np.array_equal(c['seeds'], want[idx])
        assert np.array_equal(c['axis'], axes[idx])
        # every member keeps its GLOBAL index in the Philox counter, whatever the grouping
        assert np.array_equal(c['kw']['member_index'], idx) and 'stream_offset' not in c['kw']
    assert np.array_equal(res.final_state_array()[:, 0, 0], temps)
    assert np.allclose(res.ensemble_magnetisation(), 1.0)
    with pytest.raises(TypeError):
        mp.EnsembleModel(3, base, no_such_parameter=[1, 2, 3])
    assert res.results[0].field is res.field            # all groups saw the same field: one shared array
    # members that differ in the applied field keep their own field in their per-member Results (magpy/results.py:6-92)
    def fake_amp(radius, anisotropy, axis, m0, location, Ms, alpha, T, renorm, inter, impl, dt, t_end, S, seeds, shape,
                 H0, f, tol, **kw):
        out = fake(radius, anisotropy, axis, m0, location, Ms, alpha, T, renorm, inter, impl, dt, t_end, S, seeds, shape, H0, f,
                   tol, **kw)
        out['field'] = np.full(S, H0)
        return out

    class FakePlanAmp(FakePlan):
        def __init__(self, *args, **kw):
            self.out = fake_amp(*args, **kw)
    monkeypatch.setattr(model_mod.core, 'EnsemblePlan', FakePlanAmp)
    amps = [1e3, 2e3, 1e3, 3e3, 2e3, 1e3, 1e3]
    res_a = mp.EnsembleModel(R, base, field_amplitude=amps).simulate(1e-9, 1e-12, 5, random_state=42, implicit_solve=False)
    assert [float(res_a.results[i].field[0]) for i in range(R)] == amps
    assert np.array_equal(res_a.field, np.full(5, 1e3))   # the ensemble-level field is the first member's (magpy/results.py:114)


def test_geometry_helpers():
    from magpy_b200 import geometry as g
    # the two examples of the reference's docstring (magpy/geometry/coordinates.py:80-87)
    assert np.allclose(g.chain_coordinates(3, 2.5, direction=[1, 0, 0]), [[0, 0, 0], [2.5, 0, 0], [5, 0, 0]])
    assert np.allclose(g.chain_coordinates(2, 3.0, direction=[3, 4, 0]), [[0, 0, 0], [1.8, 2.4, 0]])
    # same two draws and formula as magpy/initial_conditions.py:4-16 from the global legacy generator
    np.random.seed(11)
    theta, u = 2.0 * np.pi * np.random.rand(), np.random.rand()
    phi = np.arccos(1 - 2.0 * u)
    np.random.seed(11)
    ax = g.uniform_random_axes(5)
    assert np.allclose(ax[0], [np.sin(phi) * np.cos(theta), np.sin(phi) * np.sin(theta), np.cos(phi)])
    assert np.allclose(np.linalg.norm(ax, axis=1), 1.0)
    # BASELINE config 4 shape: 64 non-overlapping particles, reproducible from a seed
    pts = g.random_cluster_coordinates(64, 2.4e-8, rng=5)
    d = np.linalg.norm(pts[:, None] - pts[None], axis=-1) + np.eye(64)
    assert pts.shape == (64, 3) and d.min() >= 2.4e-8 and np.allclose(pts.mean(axis=0), 0, atol=1e-20)
    assert np.array_equal(pts, g.random_cluster_coordinates(64, 2.4e-8, rng=5))
    with pytest.raises(ValueError):
        g.random_cluster_coordinates(8, 1e-8, packing=0.6)


def test_shard_bounds():
    from magpy_b200.sharding import shard_bounds
    for R, G in ((10, 3), (8, 8), (1000000, 8), (5, 8), (7, 2)):
        spans = [shard_bounds(R, G, r) for r in range(G)]
        assert spans[0][0] == 0 and spans[-1][1] == R
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(hi - lo for lo, hi in spans) == -(-R // G)


WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
import torch.distributed as dist
import oracle_lib as ol
import magpy_b200 as mp
from magpy_b200 import model as model_mod

rank, world = int(sys.argv[1]), int(sys.argv[2])
dist.init_process_group('gloo', init_method='tcp://127.0.0.1:{port}', rank=rank, world_size=world)
orc = ol.load_oracle()

class GlooComm:
    # CPU stand-in for core.Comm (the library's NCCL communicator): same duck type, gloo underneath
    rank, world_size = rank, world
    def allreduce_sums(self, sums):
        import torch
        t = torch.from_numpy(sums)
        dist.all_reduce(t)
        return sums

def oracle_backed(radius, anisotropy, axis, m0, location, Ms, alpha, T, renorm, inter, impl, dt, t_end, S, seeds, shape,
                  H0, f, tol, **kw):
    # stand-in for the device call (no GPU in this test): the CPU oracle, member by member
    c = ol.make_case(N=len(radius), radius=radius, anisotropy=anisotropy, axis=np.asarray(axis).reshape(-1, 3)[:len(radius)],
                     m0=np.asarray(m0).reshape(-1, 3)[:len(radius)], location=location, Ms=Ms, alpha=alpha, T=T,
                     renorm=renorm, interactions=inter, implicit=impl, eps=tol, dt=dt, t_end=t_end, S=S,
                     field_shape=shape, H0=H0, f=f)
    tr = np.stack([ol.oracle_simulate(orc, c, seed=int(s))[2] for s in seeds])
    t, fl = ol.oracle_simulate(orc, c, seed=1)[:2]
    M = tr.sum(axis=1)
    sums = np.stack([M[:, 0].sum(0), M[:, 1].sum(0), M[:, 2].sum(0), (M[:, 2] ** 2).sum(0)], axis=1)
    return dict(N=c.N, R=len(seeds), time=t, field=fl, trajectories=tr, sums=sums, final=tr[..., -1],
                stats=dict(offset=kw.get('stream_offset'), member_index=kw.get('member_index')))

class OraclePlan:
    def __init__(self, *a, **kw):
        self.out = oracle_backed(*a, **kw)
    def run(self): pass
    def sync(self): return self.out['stats']
    def fetch(self): return dict(self.out)
model_mod.core.EnsemblePlan = OraclePlan

model_mod.core.simulate_ensemble = oracle_backed
base = mp.Model([7e-9, 7e-9], [1e5, 1e5], [[0, 0, 1.0]] * 2, [[0, 0, 1.0]] * 2, [[0, 0, 0], [0, 0, 9e-9]], 4e5, 0.1, 330.0,
                field_shape='sine', field_frequency=5e9, field_amplitude=1e4)
ens = mp.EnsembleModel(9, base)
res = ens.simulate(5e-11, 1e-12, 11, random_state=3, implicit_solve=False, shard=(rank, world), comm=GlooComm())
full = ens.simulate(5e-11, 1e-12, 11, random_state=3, implicit_solve=False)
lo, hi = mp.sharding.shard_bounds(9, world, rank)
assert res.stats[0]['offset'] == lo
assert res.final_state_array().shape[0] == hi - lo
assert np.allclose(res.final_state_array(), full.final_state_array()[lo:hi], rtol=0, atol=0)
assert np.allclose(res.ensemble_magnetisation('z'), full.ensemble_magnetisation('z'), rtol=1e-13)
assert np.allclose(res.ensemble_magnetisation_stderr(), full.ensemble_magnetisation_stderr(), rtol=1e-9)
assert np.isclose(res.energy_dissipated(), full.energy_dissipated(), rtol=1e-10)
# a sharded run over several ranks without a communicator must not silently return partial sums (ADVICE r1)
for k in ('RANK', 'WORLD_SIZE'):
    os.environ.pop(k, None)
try:
    ens.simulate(5e-11, 1e-12, 11, random_state=3, implicit_solve=False, shard=(rank, world))
    raise SystemExit('sharded run without a communicator did not raise')
except RuntimeError as exc:
    assert 'communicator' in str(exc)
# grouped override (two temperatures, interleaved) + sharding: every member keeps its global Philox index
temps = [300.0, 330.0] * 4 + [300.0]
ens2 = mp.EnsembleModel(9, base, temperature=temps)
res2 = ens2.simulate(5e-11, 1e-12, 11, random_state=3, implicit_solve=False, shard=(rank, world), comm=GlooComm())
full2 = ens2.simulate(5e-11, 1e-12, 11, random_state=3, implicit_solve=False)
seen = np.sort(np.concatenate([st['member_index'] for st in res2.stats]))
assert np.array_equal(seen, np.arange(lo, hi)), seen
assert np.array_equal(np.sort(np.concatenate([st['member_index'] for st in full2.stats])), np.arange(9))
assert np.allclose(res2.final_state_array(), full2.final_state_array()[lo:hi], rtol=0, atol=0)
assert np.allclose(res2.ensemble_magnetisation('z'), full2.ensemble_magnetisation('z'), rtol=1e-13)
dist.destroy_process_group()
print('rank', rank, 'ok')
'''


def test_sharded_ensemble_world_size_2_gloo(tmp_path):
    import socket
    from magpy_b200 import sharding  # noqa: F401
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / 'worker.py'
    script.write_text(WORKER.format(root=ROOT, tests=HERE, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r), '2'], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


@pytest.mark.parametrize('workload', ['c3', 'c4', 'c1', 'c2'])
def test_bench_reference_arm_contract(workload):
    """`bench.py --impl reference` (the reference's own CPU path on the host cores, a bounded sample per step) prints ONE
    JSON line with the contract's keys; it needs no GPU and never touches the product library."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--workload', workload,
                          '--steps', '1', '--warmup', '0'], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e', 'gpu_launches'):
        assert key in line, key
    assert line['impl'] == 'reference' and line['unit'] == 'particle-steps/s' and line['value'] > (1e5 if workload != 'c2' else 1e4)
    assert line['cpu_baseline']['kind'] in ('reference', 'port') and line['cpu_baseline']['cores'] >= 1
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['gpu_launches'] == 0
    assert line['config']['workload'].startswith(workload.upper())


def test_single_particle_radius_and_temperature_overrides_are_one_call(monkeypatch):
    """Single-particle members that differ in radius and temperature go to the device as per-member arrays in ONE call
    (radius (R, 1), temperature (R,)); for clusters the same overrides still split the ensemble by value."""
    import magpy_b200 as mp
    from magpy_b200 import model as model_mod
    calls = []

    def fake(radius, anisotropy, axis, m0, location, Ms, alpha, T, renorm, inter, impl, dt, t_end, S, seeds, shape, H0, f, tol,
             **kw):
        calls.append(dict(radius=np.array(radius), T=np.array(T), n=len(seeds)))
        n, N = len(seeds), np.asarray(anisotropy).size
        return {'N': N, 'R': n, 'time': np.arange(S) * 1.0, 'field': np.zeros(S), 'trajectories': np.zeros((n, N, 3, S)),
                'sums': np.zeros((S, 4)), 'final': np.zeros((n, N, 3)), 'stats': {}}

    class FakePlan:
        def __init__(self, *a, **kw):
            self.out = fake(*a, **kw)

        def run(self):
            pass

        def sync(self):
            return {}

        def fetch(self):
            return dict(self.out)
    monkeypatch.setattr(model_mod.core, 'simulate_ensemble', fake)
    monkeypatch.setattr(model_mod.core, 'EnsemblePlan', FakePlan)
    R = 6
    radii = [np.array([r]) for r in (5e-9, 6e-9, 7e-9, 6e-9, 8e-9, 9e-9)]
    temps = [300.0, 310.0, 300.0, 320.0, 330.0, 300.0]
    single = mp.Model([7e-9], [4e4], [[0, 0, 1.0]], [[0, 0, 1.0]], [[0, 0, 0.0]], 4e5, 0.1, 300.0)
    mp.EnsembleModel(R, single, radius=radii, temperature=temps).simulate(1e-9, 1e-12, 5, random_state=1, implicit_solve=False)
    assert len(calls) == 1 and calls[0]['n'] == R
    assert calls[0]['radius'].shape == (R, 1) and np.array_equal(calls[0]['radius'][:, 0], [r[0] for r in radii])
    assert np.array_equal(calls[0]['T'], temps)
    del calls[:]
    dimer = mp.Model([7e-9, 7e-9], [1e5, 1e5], [[0, 0, 1.0]] * 2, [[0, 0, 1.0]] * 2, [[0, 0, 0], [0, 0, 9e-9]], 4e5, 0.1, 330.0)
    mp.EnsembleModel(R, dimer, temperature=temps).simulate(1e-9, 1e-12, 5, random_state=1, implicit_solve=False)
    assert sorted(int(c['n']) for c in calls) == [1, 1, 1, 3] and all(np.ndim(c['T']) == 0 for c in calls)


def test_bench_kernel_instruction_budget():
    """K1 is bound by instruction issue: an FP64 instruction holds its sub-partition's issue port 2
    cycles, 3 with three distinct register operands, every other instruction 1.  The inner loop of the bench kernel is read
    from the SASS of the built library (no GPU needed) and held to its budget, so that a source change that makes ptxas emit
    a longer loop — or spill, or lose the sixth resident CTA — is caught here and not on the next bench run."""
    import shutil
    import subprocess
    import sys as _sys
    lib = os.path.join(ROOT, 'magpy_b200', 'libmagpy_b200.so')
    if shutil.which('cuobjdump') is None or not os.path.exists(lib):
        pytest.skip('needs cuobjdump and the built library')
    _sys.path.insert(0, os.path.join(ROOT, 'scripts'))
    import k1_cost_model as k
    import re
    for name, ins in k.functions(lib):
        if re.search('heun_single_balanced_kernelILb1ELb1ELb0', name):
            body = k.inner_loop(ins)
            cycles, hist, other = k.cost(body)
            pairs = 2                                  # two step pairs per loop trip
            fp64 = sum(hist.values())
            assert fp64 == 2 * 37 * pairs              # 37 FP64 instructions per Heun step
            assert len(body) <= 168 * pairs            # round 1: 182 per step pair
            assert cycles <= 275 * pairs               # issue-cost model; round 1: 297, measured 300.6
            res = subprocess.run(['cuobjdump', '-res-usage', lib], capture_output=True, text=True).stdout
            m = re.search(re.escape(name) + r':\s*\n\s*REG:(\d+) STACK:(\d+)', res)
            assert m and int(m.group(1)) <= 80 and int(m.group(2)) == 0      # six CTAs of 128 threads per SM, no spills
            return
    raise AssertionError('bench kernel not found in the library')
