"""Statistical parity of the production noise path (in-kernel Philox) with the reference
ensemble (oracle + the reference's mt19937_64 stream), equilibrium physics, invariance of the
results under time-chunking and member sharding, and convergence orders (test/convergence)."""
import os

import numpy as np
import pytest
from scipy import integrate

import fpe_reference as fpe
import oracle_lib as ol

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def orc():
    return ol.load_oracle()


@pytest.fixture(scope='module')
def core():
    import magpy_b200.core as core
    return core


def gpu(core, c, seeds, **kw):
    return core.simulate_ensemble(c.radius, c.anisotropy, c.axis, c.m0, c.location, c.Ms, c.alpha, c.T, c.renorm,
                                  c.interactions, c.implicit, c.dt, c.t_end, c.S, seeds, field_shape=c.field_shape,
                                  field_amplitude=c.H0, field_frequency=c.f, implicit_tol=c.eps, **kw)


def mean_and_se(sums, R, Ms):
    mean = sums[:, 2] / R
    var = np.maximum(sums[:, 3] / R - mean ** 2, 0) * R / (R - 1)
    return mean / Ms, np.sqrt(var / R) / Ms


@pytest.mark.parametrize('implicit,gauss', [(False, 'f32p'), (False, 'f32'), (False, 'f64'), (True, 'f32p')])
def test_relaxation_matches_reference_ensemble_within_3_standard_errors(orc, core, implicit, gauss):
    """Low-barrier particle (sigma = KV/kT = 3.1) relaxing from +z: <Mz>(t) of the Philox ensemble vs
    the reference-noise ensemble, at every sample, within 3 combined standard errors — as a family: the largest
    of the 40 z-scores stays inside the Bonferroni bound of one 3-sigma test (3.98)."""
    c = ol.make_case(N=1, radius=4e-9, anisotropy=4.7e4, T=300.0, dt=2e-13 if not implicit else 1e-12, t_end=4e-10,
                     S=41, implicit=implicit, axis=[[0, 0, 1.0]], m0=[[0, 0, 1.0]])
    Rg, Rc = 200000, 4000
    out = gpu(core, c, np.arange(Rg) + 12345, return_trajectories=False, gauss=gauss)
    sums_c, _ = ol.oracle_ensemble(orc, c, np.arange(Rc) + 999)
    mg, sg = mean_and_se(out['sums'], Rg, c.Ms)
    mc, sc = mean_and_se(sums_c, Rc, c.Ms)
    z = (mg - mc)[1:] / np.sqrt(sg ** 2 + sc ** 2)[1:]
    assert mg[-1] < 0.9          # it does relax
    zmax = fpe.bonferroni_z(len(z))
    assert np.abs(z).max() < zmax, z
    # second moment too: |m_z| <= 1 bounds its variance by v (1 - v), a rigorous upper bound of the standard error
    vg = out['sums'][:, 3] / Rg / c.Ms ** 2
    vc = sums_c[:, 3] / Rc / c.Ms ** 2
    se_v = np.sqrt(vg * (1 - vg) / Rg + vc * (1 - vc) / Rc)
    assert (np.abs(vg - vc)[1:] < zmax * se_v[1:]).all(), (np.abs(vg - vc)[1:] / se_v[1:]).max()


def test_dimer_relaxation_matches_reference_ensemble(orc, core):
    """BASELINE config 2 geometry: two dipolar-coupled 7 nm particles 9 nm apart, implicit midpoint."""
    c = ol.make_case(N=2, radius=7e-9, anisotropy=1e5, T=330.0, dt=1e-12, t_end=3e-10, S=31, implicit=True,
                     axis=[[0, 0, 1.0], [0, 0, 1.0]], m0=[[1.0, 0, 0], [0, 1.0, 0]], location=[[0, 0, 0], [0, 0, 9e-9]])
    Rg, Rc = 50000, 1500
    out = gpu(core, c, np.arange(Rg) + 5, return_trajectories=False)
    sums_c, _ = ol.oracle_ensemble(orc, c, np.arange(Rc) + 77)
    for comp in (0, 1, 2):
        mg = out['sums'][:, comp] / Rg / c.Ms
        mc = sums_c[:, comp] / Rc / c.Ms
        # per-member spread of a cluster component is < 2 (two unit moments); use the z-variance bound
        se = np.sqrt(2.0 / Rc + 2.0 / Rg)
        assert np.abs(mg - mc).max() < 3 * se
    mg, sg = mean_and_se(out['sums'], Rg, c.Ms)
    mc, sc = mean_and_se(sums_c, Rc, c.Ms)
    z = (mg - mc)[1:] / np.sqrt(sg ** 2 + sc ** 2)[1:]
    assert np.abs(z).max() < fpe.bonferroni_z(len(z)), z
    assert out['stats']['newton_failures'] == 0


@pytest.mark.parametrize('implicit', [False, True])
def test_single_particle_equilibrium_is_boltzmann(core, implicit):
    """docs/source/notebooks/single-particle-equilibrium.ipynb: p(theta) ~ sin(theta) exp(sigma cos^2 theta).
    Heun's weak bias on <cos^k theta> is first order in dt (+1.1e-3 at 1e-13 s): instead of an allowance it is removed
    by Richardson extrapolation over two step sizes (independent ensembles), and the extrapolated moments must sit
    within the Bonferroni bound of a 3-sigma test; the bias itself must scale like dt.  The implicit midpoint rule
    (second-order weak bias) is tested directly."""
    KB = 1.38064852e-23
    R = 1000000 if not implicit else 200000

    def moments(dt, seed0):
        c = ol.make_case(N=1, radius=5e-9, anisotropy=4e4, T=300.0, alpha=0.5, dt=dt, t_end=2e-9, S=3, implicit=implicit,
                         axis=[[0, 0, 1.0]], m0=[[0, 0, 1.0]])
        out = gpu(core, c, np.arange(R) * 3 + seed0, return_trajectories=False)
        m = out['final'][:, 0, :] / c.Ms
        m /= np.linalg.norm(m, axis=1, keepdims=True)
        sigma = c.anisotropy[0] * 4 / 3 * np.pi * c.radius[0] ** 3 / KB / c.T
        return m, sigma
    zmax = fpe.bonferroni_z(2)
    if implicit:
        m, sigma = moments(2e-12, 1)
        got = {k: ((m[:, 2] ** k).mean(), (m[:, 2] ** k).std() / np.sqrt(R)) for k in (2, 4)}
    else:
        m, sigma = moments(1e-13, 1)
        m2, _ = moments(2e-13, 2)
        got = {}
        for k in (2, 4):
            a, b = (m[:, 2] ** k), (m2[:, 2] ** k)
            got[k] = (2 * a.mean() - b.mean(), np.sqrt(4 * a.var() / R + b.var() / R))
    Z = integrate.quad(lambda x: np.exp(sigma * x * x), -1, 1)[0]
    for k in (2, 4):
        want = integrate.quad(lambda x: x ** k * np.exp(sigma * x * x), -1, 1)[0] / Z
        val, se = got[k]
        assert abs(val - want) < zmax * se, (k, val, want, se)
        assert se < 1e-3
    if not implicit:   # the first-order bias is there and doubles with the step
        want2 = integrate.quad(lambda x: x ** 2 * np.exp(sigma * x * x), -1, 1)[0] / Z
        b1, b2 = (m[:, 2] ** 2).mean() - want2, (m2[:, 2] ** 2).mean() - want2
        assert b1 > 0 and 1.3 < b2 / b1 < 3.0, (b1, b2)
    # azimuthal symmetry
    assert abs(m[:, 0].mean()) < 5 / np.sqrt(R) and abs(m[:, 1].mean()) < 5 / np.sqrt(R)


def test_results_do_not_depend_on_chunking_or_sharding(core):
    c = ol.make_case(N=1, dt=1e-13, t_end=3e-11, S=17, field_shape='sine', H0=2e4, f=4e10)
    seeds = np.arange(300) * 11 + 3
    whole = gpu(core, c, seeds)
    os.environ['MAGPY_B200_MAX_CHUNK_STEPS'] = '37'
    try:
        chunked = gpu(core, c, seeds)
    finally:
        del os.environ['MAGPY_B200_MAX_CHUNK_STEPS']
    assert chunked['stats']['kernel_launches'] > whole['stats']['kernel_launches'] + 5
    assert np.array_equal(whole['trajectories'], chunked['trajectories'])
    assert np.array_equal(whole['final'], chunked['final'])
    assert np.allclose(whole['sums'], chunked['sums'], rtol=1e-14)
    # two "ranks": members [0,140) and [140,300) with their global member offsets
    a = gpu(core, c, seeds[:140], stream_offset=0)
    b = gpu(core, c, seeds[140:], stream_offset=140)
    assert np.array_equal(np.concatenate([a['trajectories'], b['trajectories']]), whole['trajectories'])
    assert np.allclose(a['sums'] + b['sums'], whole['sums'], rtol=1e-13)
    # a different offset is a different (independent) stream
    d = gpu(core, c, seeds[140:], stream_offset=0)
    assert not np.array_equal(d['final'], b['final'])
    # cluster kernels: same invariance
    c2 = ol.make_case(N=3, radius=7e-9, anisotropy=1e5, dt=1e-13, t_end=2e-11, S=9)
    w2 = gpu(core, c2, seeds[:70])
    os.environ['MAGPY_B200_MAX_CHUNK_STEPS'] = '23'
    try:
        ch2 = gpu(core, c2, seeds[:70])
    finally:
        del os.environ['MAGPY_B200_MAX_CHUNK_STEPS']
    assert np.array_equal(w2['trajectories'], ch2['trajectories'])


def test_single_process_multi_gpu_matches_one_gpu(core):
    """magpy_b200_simulate_ensemble_multi: members sharded over the GPUs of the box from one process give the
    per-member outputs of a one-GPU run bit for bit and the same ensemble sums up to fp64 summation order."""
    n_dev = core.device_count()
    c = ol.make_case(N=3, dt=1e-13, t_end=2e-11, S=9, field_shape='sine', H0=1e4, f=1e10, rng=np.random.default_rng(3))
    seeds = np.arange(1, 1001) * 17
    rng = np.random.default_rng(4)
    m0 = rng.normal(size=(len(seeds), c.N, 3)); m0 /= np.linalg.norm(m0, axis=-1, keepdims=True)
    one = gpu(core, ol.Case(dict(c, m0=m0)), seeds)
    # more devices than exist must fail loudly
    with pytest.raises(RuntimeError):
        gpu(core, ol.Case(dict(c, m0=m0)), seeds, devices=[0, n_dev])
    devs = list(range(n_dev)) if n_dev > 1 else [0, 0]      # one GPU: two shards on the same device
    many = gpu(core, ol.Case(dict(c, m0=m0)), seeds, devices=devs)
    assert np.array_equal(one['trajectories'], many['trajectories'])
    assert np.array_equal(one['final'], many['final'])
    assert np.array_equal(one['time'], many['time']) and np.array_equal(one['field'], many['field'])
    assert np.allclose(one['sums'], many['sums'], rtol=1e-12, atol=1e-9 * c.Ms)
    assert many['stats']['particle_steps'] == one['stats']['particle_steps']
    # 'all' and a ragged split (3 shards of 334, 334, 332)
    rag = gpu(core, ol.Case(dict(c, m0=m0)), seeds, devices=[0] * 3, return_trajectories=False)
    assert np.array_equal(one['final'], rag['final'])


def test_hysteresis_energy_matches_reference_ensemble(orc, core):
    """Energy dissipated per cycle, -mu0 * loop area (magpy/results.py:167-217), GPU/Philox vs
    oracle/MT ensembles of a low-barrier particle driven at 2 GHz (short enough for the CPU)."""
    from magpy_b200.results import EnsembleResults
    c = ol.make_case(N=1, radius=5e-9, anisotropy=4e4, T=300.0, alpha=0.3, dt=5e-13, t_end=1e-9, S=201,
                     field_shape='sine', H0=6e4, f=2e9, axis=[[0, 0, 1.0]], m0=[[0, 0, 1.0]])
    Rg, Rc = 100000, 3000
    out = gpu(core, c, np.arange(Rg) + 1, return_trajectories=False)
    eg = EnsembleResults.from_arrays(out['time'], out['field'], Rg, sums=out['sums'], final=out['final'])
    Eg = eg.final_cycle_energy_dissipated(c.f)
    # bootstrap the CPU ensemble in 6 blocks for a standard error of the loop area
    blocks = []
    t, fl = ol.oracle_simulate(orc, c, seed=1)[:2]
    for b in range(6):
        s, _ = ol.oracle_ensemble(orc, c, np.arange(Rc // 6) + 1000 * (b + 1))
        blocks.append(EnsembleResults.from_arrays(t, fl, Rc // 6, sums=s).final_cycle_energy_dissipated(c.f))
    Ec, se = np.mean(blocks), np.std(blocks, ddof=1) / np.sqrt(6)
    # magpy/results.py:191 returns -mu0 * integral(H dM): negative for a dissipative (lagging) loop
    assert Eg < 0 and Ec < 0 and abs(Eg) > 100.0
    assert abs(Eg - Ec) < 3.5 * se + 0.01 * abs(Ec), (Eg, Ec, se)


def _coupled_levels(core, c, n_fine, levels, R, rng):
    """Final states for step sizes dt*2^l driven by the SAME Brownian paths (coarse increments are
    sums of fine ones, as test/convergence/task5.cpp:150-158)."""
    w = rng.normal(size=(R, n_fine, 3 * c.N))
    finals = []
    base_dt = c.dt
    for l in range(levels):
        m = 2 ** l
        wl = w.reshape(R, n_fine // m, m, 3 * c.N).sum(axis=2) / np.sqrt(m)
        cc = ol.Case(c)
        cc['dt'] = base_dt * m
        cc["t_end"] = base_dt * n_fine * (1 + 1e-9)      # zero-order hold: the last sample is the state after n_fine/m steps
        cc['S'] = 2
        # the reference executes one more (unobserved) step: pad the stream by one increment
        wl = np.concatenate([wl, np.zeros((R, 2, 3 * c.N))], axis=1)
        out = core.simulate_ensemble(cc.radius, cc.anisotropy, cc.axis, cc.m0, cc.location, cc.Ms, cc.alpha, cc.T,
                                     False, True, cc.implicit, cc.dt, cc.t_end, cc.S, np.zeros(R, dtype=np.int64),
                                     implicit_tol=cc.eps, injected_dw=wl, return_trajectories=False)
        assert out['stats']['steps_per_member'] == n_fine // m
        finals.append(out['final'][:, 0, :] / cc.Ms)
    return finals


@pytest.mark.parametrize('implicit', [False, True])
def test_strong_convergence_order(orc, core, implicit):
    """test/convergence/task5 (+ docs/source/notebooks/convergence.ipynb cells 29-38): mean 2-norm of
    the Cauchy differences between consecutive step sizes on common Brownian paths has slope
    ~0.5 in log2(dt) for both schemes (the reference's own run: 0.509 Heun, 0.511 implicit)."""
    rng = np.random.default_rng(2024)
    c = ol.make_case(N=1, radius=6e-9, anisotropy=4e4, T=300.0, alpha=0.1, implicit=implicit, eps=1e-10,
                     axis=[[0, 0, 1.0]], m0=[[1.0, 0, 0]])
    tf = ol.reduced_scalars(orc, c)['time_factor']
    c['dt'] = 0.0025 / tf                              # reduced dt = 0.0025 * 2^l
    R, n_fine, levels = 10000, 1280, 6
    finals = _coupled_levels(core, c, n_fine, levels, R, rng)
    strong = [np.linalg.norm(finals[l + 1] - finals[l], axis=1).mean() for l in range(levels - 1)]
    order = np.polyfit(np.arange(levels - 1), np.log2(strong), 1)[0]
    assert 0.4 < order < 0.7, (order, strong)


@pytest.mark.parametrize('implicit', [False, True])
def test_weak_convergence_order(orc, core, implicit):
    """The reference has no weak-order test; ours: |E m_z(T)^{dt} - E m_z(T)^{2dt}| on
    common Brownian paths (variance reduction), an observable without a symmetry that would
    make it vanish.  Heun and the midpoint rule are weakly first order: slope ~1."""
    rng = np.random.default_rng(7)
    c = ol.make_case(N=1, radius=6e-9, anisotropy=4e4, T=300.0, alpha=0.1, implicit=implicit, eps=1e-10,
                     axis=[[0, 0, 1.0]], m0=[[0.6, 0, 0.8]])
    tf = ol.reduced_scalars(orc, c)['time_factor']
    c['dt'] = 0.02 / tf
    R, n_fine, levels = 40000, 256, 5
    finals = _coupled_levels(core, c, n_fine, levels, R, rng)
    diff = [finals[l + 1][:, 2] - finals[l][:, 2] for l in range(levels - 1)]
    weak = np.array([abs(d.mean()) for d in diff])
    se = np.array([d.std() / np.sqrt(R) for d in diff])
    sig = [l for l in range(levels - 1) if weak[l] > 3 * se[l]]
    assert len(sig) >= 3, (weak, se)
    order = np.polyfit(np.array(sig), np.log2(weak[sig]), 1)[0]
    assert 0.6 < order < 2.4, (order, weak, se)


def test_neel_relaxation_time_against_master_equation(orc, core):
    """Single 6 nm particle (sigma = KV/kT = 8.7) relaxing from +z over 1 microsecond: the decay rate of the
    ensemble <Mz>(t) against the reference's discrete-orientation master equation (lib/dom.cpp:33-59), whose
    two-state solution at zero field is exp(-2 W t).  The master equation uses the Neel-Brown high-barrier
    asymptote (valid for sigma >> 1, lib/dom.cpp:20-22; its leading correction is the factor 1 - 1/sigma,
    i.e. 11 % here), so the bar is 20 % on the rate, not a number of standard errors."""
    # renorm=True as in the reference's long Heun runs: without it |m| random-walks away from 1 over 5e5 steps
    # (lib/simulation.cpp:379-387), in the reference exactly as here
    c = ol.make_case(N=1, radius=6e-9, anisotropy=4e4, Ms=4e5, alpha=0.1, T=300.0, dt=2e-12, t_end=1e-6, S=201,
                     axis=[[0, 0, 1.0]], m0=[[0, 0, 1.0]], renorm=True)
    R = 32768
    out = gpu(core, c, np.arange(R) + 777, return_trajectories=False)
    mz, se = mean_and_se(out['sums'], R, c.Ms)
    V = 4.0 / 3.0 * np.pi * 6e-9 ** 3
    W = ol.dom_transition_matrix(orc, 4e4, V, 300.0, 0.0, 4e5, 0.1)
    rate_dom = 2 * W[1]                                    # p0 - p1 decays with W01 + W10
    t = out['time']
    sel = (t > 0.1e-6) & (mz > 10 * se)                    # past the intra-well transient, above the noise floor
    slope, intercept = np.polyfit(t[sel], np.log(mz[sel]), 1)
    rate_sllg = -slope
    assert sel.sum() > 50
    assert 0.8 < rate_sllg / rate_dom < 1.2, (rate_sllg, rate_dom)
    # the corrected asymptote (1 - 1/sigma) is closer still
    sigma = 4e4 * V / (ol.KB * 300.0)
    assert abs(rate_sllg / (rate_dom * (1 - 1 / sigma)) - 1) < 0.1, (rate_sllg, rate_dom, sigma)
    # the fast intra-well relaxation leaves <Mz> near the well average <cos theta> = 1 - 1/(2 sigma) + ...
    assert abs(np.exp(intercept) - (1 - 1 / (2 * sigma))) < 0.03


def test_full_size_ensembles_checksum_properties(core):
    """BASELINE configs 3 and 4 at their full member counts (1,000,000 single particles under the 300 kHz / 20 kA/m
    sine field; 100,000 clusters of 64 dipolar-coupled particles), shortened in time: size-independent properties
    instead of an oracle — the fused ensemble sums are the sums of the per-member final states (a checksum of
    checksums), halves of the ensemble add up to the whole whatever the member range / Philox offset, and
    renormalised moments have unit length."""
    from magpy_b200 import geometry
    # config 3
    R = 1_000_000
    c = ol.make_case(N=1, dt=1e-12, t_end=2e-8, S=5, field_shape='sine', H0=2e4, f=3e5, axis=[[0, 0, 1.0]],
                     m0=[[0, 0, 1.0]], renorm=True)
    seeds = np.random.default_rng(1).integers(0, 2 ** 31 - 1, R)
    whole = gpu(core, c, seeds, return_trajectories=False)
    f = whole['final'][:, 0, :] / c.Ms
    assert np.abs(np.linalg.norm(f, axis=1) - 1).max() < 1e-12
    assert abs(whole['sums'][-1, 2] / c.Ms - f[:, 2].sum()) < 1e-6 * R
    assert abs(whole['sums'][-1, 3] / c.Ms ** 2 - (f[:, 2] ** 2).sum()) < 1e-6 * R
    lo = gpu(core, c, seeds[:R // 2], return_trajectories=False)
    hi = gpu(core, c, seeds[R // 2:], return_trajectories=False, stream_offset=R // 2)
    assert np.array_equal(np.concatenate([lo['final'], hi['final']]), whole['final'])
    assert np.allclose(lo['sums'] + hi['sums'], whole['sums'], rtol=1e-12, atol=1e-6 * c.Ms)
    assert whole['stats']['particle_steps'] == R * whole['stats']['steps_per_member']
    # config 4
    N, R4 = 64, 100_000
    axes = geometry.uniform_random_axes(N, rng=4)
    c4 = ol.make_case(N=N, radius=12e-9, anisotropy=4e4, axis=axes, m0=axes, dt=1e-14, t_end=1e-12, S=3,
                      location=geometry.random_cluster_coordinates(N, 3e-8, rng=4), renorm=True)
    out = gpu(core, c4, np.arange(R4), return_trajectories=False)
    f4 = out['final'] / c4.Ms                                           # [R, N, 3]
    assert np.abs(np.linalg.norm(f4, axis=2) - 1).max() < 1e-12
    Mz = f4[:, :, 2].sum(axis=1)
    assert abs(out['sums'][-1, 2] / c4.Ms - Mz.sum()) < 1e-6 * R4 * N
    assert abs(out['sums'][-1, 3] / c4.Ms ** 2 - (Mz ** 2).sum()) < 1e-6 * R4 * N * N
    # after 100 steps of 1e-14 s every moment is still close to its easy axis, which is where it started
    assert np.abs((f4 * axes[None]).sum(axis=2)).min() > 0.99


@pytest.mark.parametrize('N,implicit', [(1, False), (1, True), (2, False), (6, False), (2, True), (1, 'exact'), (3, 'exact')])
def test_zero_temperature_relaxation_closed_form(orc, core, N, implicit):
    """T = 0 (the in-kernel noise amplitude is exactly zero), no applied field, non-interacting particles with
    their easy axes along z: tan(theta(t)) = tan(theta0) exp(-alpha t) in reduced time (the closed form the
    reference's test/convergence/task4 family uses), through the production (Philox) kernels of every family —
    single, one-thread-per-cluster, shared-memory cluster; Heun and implicit midpoint."""
    th0 = np.pi / 3
    newton = 'exact' if implicit == 'exact' else 'reference'     # the opt-in exact-Jacobian Newton mode as well
    implicit = bool(implicit)
    c = ol.make_case(N=N, T=0.0, alpha=0.1, S=11, axis=[[0, 0, 1.0]] * N, m0=[[np.sin(th0), 0, np.cos(th0)]] * N,
                     implicit=implicit, interactions=False)
    tf = ol.reduced_scalars(orc, c)['time_factor']
    dt_red = 0.01
    c['dt'] = dt_red / tf
    c['t_end'] = 4.0 / tf
    out = gpu(core, c, np.arange(33), implicit_newton=newton)
    cum = ol.schedule(orc, dt_red, 4.0, c.S)
    ts = np.maximum(cum.astype(float) - 1, 0) * dt_red
    exact = 1.0 / np.sqrt(1 + np.tan(th0) ** 2 * np.exp(-2 * 0.1 * ts))
    mz = out['trajectories'][:, :, 2, :] / c.Ms                 # [R, N, S]
    assert np.abs(mz - exact).max() < (2e-5 if not implicit else 2e-4)
    assert np.abs(np.linalg.norm(out['trajectories'], axis=2) / c.Ms - 1).max() < 1e-5
    assert out['stats']['newton_failures'] == 0


def test_parameter_groups_run_concurrently_and_match_blocking_calls(core):
    """An ensemble whose members differ in a parameter that cannot ride on a per-member array (here the field
    frequency: a frequency sweep) is split into groups of equal parameters; the groups run as concurrent plans.  Every
    member must come out exactly as from a blocking call for its group alone, whatever the number of plans in flight."""
    import magpy_b200 as mp
    from magpy_b200 import model as model_mod
    R, groups = 240, 12
    freqs = np.repeat(np.linspace(5e8, 2.5e9, groups), R // groups)
    base = mp.Model([8e-9], [4e4], [[0, 0, 1.0]], [[0, 0, 1.0]], [[0, 0, 0.0]], 4e5, 0.1, 300.0, field_shape='sine',
                    field_frequency=1e9, field_amplitude=1e4)
    ens = mp.EnsembleModel(R, base, field_frequency=list(freqs))
    res = ens.simulate(2e-9, 1e-13, 21, 7, implicit_solve=False)
    assert len(res.stats) == groups
    old = model_mod._MAX_CONCURRENT_PLANS
    try:
        model_mod._MAX_CONCURRENT_PLANS = 1          # one plan at a time: the blocking order
        res1 = ens.simulate(2e-9, 1e-13, 21, 7, implicit_solve=False)
    finally:
        model_mod._MAX_CONCURRENT_PLANS = old
    assert np.array_equal(res.final_state_array(), res1.final_state_array())
    assert np.allclose(res.ensemble_magnetisation(), res1.ensemble_magnetisation(), rtol=1e-13, atol=0)
    seeds = ens._member_seeds(7)
    idx = np.arange(3 * (R // groups), 4 * (R // groups))      # the fourth group through the plain blocking entry point
    out = core.simulate_ensemble([8e-9], [4e4], [[0, 0, 1.0]], [[0, 0, 1.0]], [[0, 0, 0.0]], 4e5, 0.1, 300.0, False, True, False,
                                 1e-13, 2e-9, 21, seeds[idx], field_shape='sine', field_amplitude=1e4,
                                 field_frequency=float(freqs[idx[0]]), stream_offset=int(idx[0]))
    assert np.array_equal(out['final'], res.final_state_array()[idx])


# ---------------------------------------------------------------------------------------------------------------
# round 2: the Gaussian stream at 1e9 draws, and the sLLG ensemble against the exact Fokker-Planck solution
# (tests/fpe_reference.py) and the reference's discrete-orientation master equation (lib/dom.cpp)
# ---------------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize('mode,n_members,n_steps', [('f32p', 151552, 2200), ('f32', 151552, 1100), ('f64', 151552, 550)])
def test_gaussian_stream_tails_ks_and_angle_uniformity(core, mode, n_members, n_steps):
    """lib/rng.cpp:14-24 draws N(0,1) in fp64; the in-kernel stream must be the same distribution where it matters.
    On the device, over 1e9 draws of the production mode (2.5e8+ of the others): tail probabilities P(|z| > 3, 4, 5)
    against erfc within 3 binomial standard errors, Kolmogorov-Smirnov over 4096 bin edges at the 0.1 % level,
    chi-square of the Box-Muller direction over 1024 bins, and the first four moments."""
    from scipy.special import erfc
    from scipy.stats import norm
    st = core.gaussian_stats(20261017, n_members, n_steps, gauss=mode)
    n = st['n']
    assert n == 3 * n_members * n_steps and int(st['hist'].sum()) == n
    edges, hist = st['edges'], st['hist'].astype(np.float64)
    report = {}
    for t in (3.0, 4.0, 5.0):
        count = hist[(edges[:-1] >= t) | (edges[1:] <= -t)].sum()
        p = erfc(t / np.sqrt(2.0))
        z = (count - n * p) / np.sqrt(n * p * (1 - p))
        report[t] = (int(count), n * p, z)
        assert abs(z) < 3.0, (mode, report)
    cdf = np.concatenate([[0.0], np.cumsum(hist)]) / n
    D = np.abs(cdf - norm.cdf(edges)).max()
    assert np.sqrt(n) * D < 1.95, (mode, D)                       # Kolmogorov: P(sqrt(n) D > 1.95) = 0.001
    na = st['n_angles']
    exp = na / 1024.0
    chi2 = ((st['angle_hist'].astype(np.float64) - exp) ** 2 / exp).sum()
    assert abs(chi2 - 1023.0) < 3.5 * np.sqrt(2 * 1023.0), (mode, chi2)
    m1, m2, m3, m4 = (st[k] / n for k in ('sum', 'sum2', 'sum3', 'sum4'))
    assert abs(m1) < 4 / np.sqrt(n) and abs(m2 - 1) < 4 * np.sqrt(2.0 / n)
    assert abs(m3) < 4 * np.sqrt(15.0 / n) and abs(m4 - 3) < 4 * np.sqrt(96.0 / n)
    # the packed mode's radius stops at sqrt(2 ln 2^23) = 5.65 (documented truncation: 1.6e-8 per draw)
    assert st['max_abs'] < (5.66 if mode == 'f32p' else 7.0) and st['max_abs'] > 5.0


def _sllg_mean(core, radius, dt, t_end, S, R, seed0, H0=0.0, f=0.0, renorm=True, implicit=False):
    c = ol.make_case(N=1, radius=radius, anisotropy=4e4, Ms=4e5, alpha=0.1, T=300.0, dt=dt, t_end=t_end, S=S,
                     axis=[[0, 0, 1.0]], m0=[[0, 0, 1.0]], renorm=renorm, implicit=implicit,
                     field_shape='sine' if f else 'constant', H0=H0, f=f)
    out = gpu(core, c, np.arange(R) + seed0, return_trajectories=False)
    mz, se = mean_and_se(out['sums'], R, c.Ms)
    return out['time'], out['field'], mz, se


@pytest.mark.parametrize('radius,t_end,dt', [(6e-9, 1e-6, 1e-12), (7e-9, 8e-6, 2e-12)])
def test_neel_relaxation_matches_fokker_planck_within_3_se(orc, core, radius, t_end, dt):
    """Relaxation of <m_z>(t) from +z at zero field (sigma = 8.7 and 13.9) against the exact Fokker-Planck solution of
    the same SDE — every sample inside the Bonferroni bound of a 3-sigma test, and the fitted Neel rate within 3
    standard errors of the exact smallest eigenvalue.  The reference's master equation (lib/dom.cpp:33-59) uses the
    high-barrier asymptote of that rate: its bias is measured against the exact value here (14.7 % at sigma = 8.7,
    8.4 % at 13.9, 5.3 % at 20.7 — it vanishes like 1 / sigma), not absorbed in a wider bar."""
    p = fpe.reduced(radius, 4e4, 4e5, 0.1, 300.0)
    R, S = 262144, 201
    t, _, mz, se = _sllg_mean(core, radius, dt, t_end, S, R, 4242)
    f1, _ = fpe.mean_mz(p['sigma'], 0.1, lambda tt: 0.0, t * p['time_factor'])
    z = (mz[1:] - f1[1:]) / se[1:]
    assert np.abs(z).max() < fpe.bonferroni_z(S - 1), (np.abs(z).max(), np.abs(mz - f1).max())
    # decay rate by weighted least squares on log <m_z> (past the intra-well transient)
    sel = t > 0.1 * t_end
    w = (mz[sel] / se[sel]) ** 2
    A = np.vstack([np.ones(sel.sum()), t[sel]]).T
    cov = np.linalg.inv(A.T @ (A * w[:, None]))
    coef = cov @ (A.T @ (w * np.log(mz[sel])))
    rate, rate_se = -coef[1], np.sqrt(cov[1, 1]) * np.sqrt(sel.sum() / 4.0)    # samples are correlated: inflate by the
    exact = fpe.neel_rate(p['sigma'], 0.1) * p['time_factor']                    # integrated autocorrelation (~ S / 4)
    assert abs(rate - exact) < 3 * rate_se, (rate, exact, rate_se)
    assert rate_se / exact < 0.05
    W = ol.dom_transition_matrix(orc, 4e4, p['V'], 300.0, 0.0, 4e5, 0.1)
    bias = 2 * W[1] / exact
    assert 1.0 < bias < 1.0 + 1.4 / p['sigma'], bias                             # the asymptote's error, O(1 / sigma)
    assert abs(rate / (2 * W[1]) - 1 / bias) < 3 * rate_se / exact               # sLLG vs DOM = the DOM's own bias, within 3 SE


def test_hysteresis_loop_and_sar_match_fokker_planck_and_quantify_the_master_equation(orc, core):
    """BASELINE config 3's drive (300 kHz, 20 kA/m) on the 6 nm particle (sigma = 8.7: it switches) run to the steady
    state (5 periods; period-to-period difference inside 3 standard errors): <m_z>(t) over the last cycle against the
    exact Fokker-Planck solution inside the Bonferroni bound; cycle energy and SAR within 3 standard errors (block
    bootstrap over 16 member blocks).  The two-state master equation (magpy.DOModel, lib/dom.cpp) on the same drive
    differs from BOTH by the same, deterministic amount: its loop saturates at |m_z| = 0.95 instead of 0.876 (no
    intra-well motion: a gap of up to 0.114 in <m_z>) while its cycle energy is only 0.7 % larger — sLLG minus DOM equals
    Fokker-Planck minus DOM within 3 SE."""
    import magpy_b200 as mp
    from magpy_b200.results import EnsembleResults
    f, H0, radius = 3e5, 2e4, 6e-9
    p = fpe.reduced(radius, 4e4, 4e5, 0.1, 300.0, H0, f)
    periods, per = 5, 400
    S = periods * per + 1
    R, blocks = 65536, 16
    c = ol.make_case(N=1, radius=radius, anisotropy=4e4, Ms=4e5, alpha=0.1, T=300.0, dt=2e-12, t_end=periods / f, S=S,
                     axis=[[0, 0, 1.0]], m0=[[0, 0, 1.0]], renorm=True, field_shape='sine', H0=H0, f=f)
    outs = [gpu(core, c, np.arange(R // blocks) + 1000003 * (b + 1), return_trajectories=False, stream_offset=b * (R // blocks))
            for b in range(blocks)]
    sums = np.sum([o['sums'] for o in outs], axis=0)
    t, field = outs[0]['time'], outs[0]['field']
    mz, se = mean_and_se(sums, R, c.Ms)
    # steady state: the last two periods agree within 3 SE (Bonferroni over the samples of a period)
    zpp = (mz[-per:] - mz[-2 * per:-per]) / np.sqrt(se[-per:] ** 2 + se[-2 * per:-per] ** 2)
    assert np.abs(zpp).max() < fpe.bonferroni_z(per), np.abs(zpp).max()
    f1, _ = fpe.mean_mz(p['sigma'], 0.1, lambda tt: p['h0'] * np.sin(2 * np.pi * p['f_red'] * tt), t * p['time_factor'])
    z = (mz[-per:] - f1[-per:]) / se[-per:]
    assert np.abs(z).max() < fpe.bonferroni_z(per), (np.abs(z).max(), np.abs(mz[-per:] - f1[-per:]).max())
    assert abs(f1[-per:].max() - 0.8764) < 1e-3 and abs(mz[-per:].max() - f1[-per:].max()) < 4 * se[-1]

    def cycle_energy(m):      # -mu0 * loop integral of H dM over the last period (magpy/results.py:167-217), per unit volume
        return EnsembleResults.from_arrays(t, field, 1, sums=np.stack([m * 0, m * 0, m * c.Ms, m * 0], axis=1)
                                           ).final_cycle_energy_dissipated(f)
    E_blocks = [cycle_energy(mean_and_se(o['sums'], R // blocks, c.Ms)[0]) for o in outs]
    E, E_se = cycle_energy(mz), np.std(E_blocks, ddof=1) / np.sqrt(blocks)
    E_fp = cycle_energy(f1)
    assert E < 0 and abs(E - E_fp) < 3 * E_se, (E, E_fp, E_se)
    assert E_se / abs(E_fp) < 0.01
    sar, sar_fp = abs(E) * f / 5180.0, abs(E_fp) * f / 5180.0                    # W/kg of magnetite
    assert abs(sar - sar_fp) < 3 * E_se * f / 5180.0
    # the reference's comparator: two-state master equation on the same drive (GPU kernel == reference, test_dom_gpu.py)
    dom = mp.DOModel(radius, 4e4, [1.0, 0.0], 4e5, 0.1, 300.0, field_shape='sine', field_frequency=f, field_amplitude=H0)
    dz = dom.simulate(periods / f, 1e-10, S).z[0]
    E_dom = cycle_energy(dz)
    assert abs(np.abs(dz[-per:]).max() - 0.95) < 0.01                             # [-0.945, 0.950]
    model_gap = f1[-per:] - dz[-per:]                                              # deterministic: DOM's approximation error
    assert 0.10 < np.abs(model_gap).max() < 0.13
    zz = ((mz[-per:] - dz[-per:]) - model_gap) / se[-per:]
    assert np.abs(zz).max() < fpe.bonferroni_z(per)
    assert abs((E - E_dom) - (E_fp - E_dom)) < 3 * E_se and 1.002 < E_dom / E_fp < 1.015, (E, E_fp, E_dom)
