"""GPU parity: the CUDA path (through magpy_b200.core -> C ABI) against the CPU oracle on the
same inputs and the same Wiener increments.  Bar (north_star): 1e-10 relative in fp64."""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu

TOL = 1e-10


@pytest.fixture(scope='module')
def orc():
    return ol.load_oracle()


@pytest.fixture(scope='module')
def core():
    import magpy_b200.core as core
    return core


def gpu_run(core, c, seeds, dW=None, axis=None, m0=None, **kw):
    return core.simulate_ensemble(
        c.radius, c.anisotropy, c.axis if axis is None else axis, c.m0 if m0 is None else m0, c.location,
        c.Ms, c.alpha, c.T, c.renorm, c.interactions, c.implicit, c.dt, c.t_end, c.S, seeds,
        field_shape=c.field_shape, field_amplitude=c.H0, field_frequency=c.f, implicit_tol=c.eps,
        injected_dw=dW, **kw)


def injected_pair(orc, core, c, seeds, per_member=False):
    """Oracle vs GPU with the reference's mt19937_64/normal stream of each member's seed injected."""
    n_steps = ol.steps_executed(orc, c)
    R = len(seeds)
    rng = np.random.default_rng(7)
    axis = m0 = None
    if per_member:
        def unit(v):
            return v / np.linalg.norm(v, axis=-1, keepdims=True)
        axis = unit(rng.normal(size=(R, c.N, 3)))
        m0 = unit(rng.normal(size=(R, c.N, 3)))
    dW = np.stack([ol.mt_normal(orc, int(s), n_steps * 3 * c.N).reshape(n_steps, 3 * c.N) for s in seeds])
    ref = []
    newton = []
    for i, s in enumerate(seeds):
        t, fl, m, it, fails = ol.oracle_simulate(orc, c, seed=int(s), axis=None if axis is None else axis[i],
                                                 m0=None if m0 is None else m0[i])
        ref.append(m)
        newton.append(it)
    ref = np.stack(ref)
    out = gpu_run(core, c, seeds, dW=dW, axis=axis, m0=m0)
    return t, fl, ref, out, newton


def assert_traj(ref, out, c):
    err = np.abs(out['trajectories'] - ref).max() / c.Ms
    assert err <= TOL, 'max |dm|/Ms = %.3e' % err
    # fused ensemble sums and final states are consistent with the trajectories
    M = out['trajectories'].sum(axis=1)                      # [R,3,S]
    sums = np.stack([M[:, 0].sum(0), M[:, 1].sum(0), M[:, 2].sum(0), (M[:, 2] ** 2).sum(0)], axis=1)
    scale = np.array([c.Ms, c.Ms, c.Ms, c.Ms ** 2]) * len(ref) * c.N
    assert np.abs(out['sums'] - sums).max() / scale.max() < 1e-12
    assert np.abs((out['sums'] - sums) / scale).max() < 1e-12
    assert np.array_equal(out['final'], out['trajectories'][..., -1])


def test_philox_words_match_random123_and_oracle(orc, core):
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        assert tuple(core.philox_words(ctr, key)) == want
        assert tuple(ol.philox(orc, ctr, key)) == want
    rng = np.random.default_rng(3)
    for _ in range(20):
        ctr = [int(x) for x in rng.integers(0, 2 ** 32, 4)]
        key = [int(x) for x in rng.integers(0, 2 ** 32, 2)]
        assert core.philox_words(ctr, key) == ol.philox(orc, ctr, key)


@pytest.mark.parametrize('mode,tol', [('f64', 1e-13), ('f32', 2e-5), ('f32p', 2e-5)])
def test_gaussian_stream_matches_oracle(orc, core, mode, tol):
    seed, member, particle, first, n = 123456789012345, 77, 3, 1000, 4096
    got = core.gaussians(seed, member, particle, first, n, gauss=mode)
    want = np.array([ol.philox_gauss3(orc, seed, member, particle, first + i, {'f32': 0, 'f64': 1, 'f32p': 2}[mode])
                     for i in range(n)])
    assert np.abs(got - want).max() < tol
    # sanity of the stream itself
    big = core.gaussians(seed, 1, 0, 1, 400000, gauss=mode).ravel()
    assert abs(big.mean()) < 5 / np.sqrt(big.size)
    assert abs(big.var() - 1) < 5 * np.sqrt(2 / big.size)
    assert abs((big ** 4).mean() - 3) < 0.05


@pytest.mark.parametrize('field_shape,H0,f,renorm', [
    ('constant', 0.0, 0.0, False), ('constant', 3e4, 0.0, True), ('sine', 2e4, 3e9, False), ('square', 2e4, 5e9, False)])
def test_heun_single_injected(orc, core, field_shape, H0, f, renorm):
    c = ol.make_case(N=1, dt=1e-14, t_end=2e-11, S=64, field_shape=field_shape, H0=H0, f=f, renorm=renorm,
                     axis=[[0, 0, 1.0]], m0=[[1.0, 0, 0]])
    seeds = np.array([1001, 5, 77, 123456, 9, 31337, 2, 42, 8, 100])
    t, fl, ref, out, _ = injected_pair(orc, core, c, seeds)
    assert_traj(ref, out, c)
    assert np.allclose(out['time'], t, rtol=0, atol=0) and np.allclose(out['field'], fl, rtol=1e-15, atol=0)


def test_heun_single_config1_full_length(orc, core):
    """BASELINE config 1: 12 nm particle, Heun dt=1e-14 to 1e-9 s (1e5 steps), pathwise."""
    c = ol.make_case(N=1, dt=1e-14, t_end=1e-9, S=1000, axis=[[0, 0, 1.0]], m0=[[1.0, 0, 0]])
    seeds = np.array([1001, 2002, 3003, 4004])
    t, fl, ref, out, _ = injected_pair(orc, core, c, seeds)
    assert_traj(ref, out, c)


def test_heun_single_per_member_axes_and_ragged_block(orc, core):
    c = ol.make_case(N=1, dt=2e-14, t_end=1e-11, S=33, field_shape='sine', H0=1e4, f=1e10)
    seeds = np.arange(1, 131 + 1) * 17            # 131 members: not a multiple of the CTA size
    t, fl, ref, out, _ = injected_pair(orc, core, c, seeds, per_member=True)
    assert_traj(ref, out, c)


@pytest.mark.parametrize('axis', [[0, 0, 1.0], [0.6, 0, 0.8]])        # easy axis along z: specialised kernel
@pytest.mark.parametrize('field_shape,H0,f,renorm,eps', [
    ('constant', 0.0, 0.0, False, 1e-9), ('sine', 2e4, 3e9, False, 1e-9), ('constant', 1e4, 0.0, True, 1e-6),
    # strong field (h = H/H_k = 3.1): the quasi-Newton matrix I - a'/2 - B'.w/2 is far from the identity
    ('constant', 5e5, 0.0, False, 1e-9)])
def test_implicit_single_injected(orc, core, field_shape, H0, f, renorm, eps, axis):
    c = ol.make_case(N=1, dt=1e-13, t_end=5e-11, S=40, implicit=True, eps=eps, field_shape=field_shape, H0=H0, f=f,
                     renorm=renorm, axis=[axis], m0=[[1.0, 0, 0]])
    seeds = np.array([1001, 5, 77, 123456, 9, 31337])
    t, fl, ref, out, newton = injected_pair(orc, core, c, seeds)
    assert_traj(ref, out, c)
    # the quasi-Newton iteration is reproduced, not just its fixed point: same iteration counts
    # (the oracle also runs the reference's last, never-sampled step; the product skips it)
    assert out['stats']['newton_iterations'] == sum(n[0] - n[2] for n in newton)
    assert out['stats']['newton_max_iterations'] <= max(n[1] for n in newton)
    assert out['stats']['newton_failures'] == 0


# Scalar kernels.  N = 2..7: one thread per cluster (small_heun.cu); 8..32: two particles per thread, pair table in
# shared memory; 40: four per thread; 64: four per thread with ONE moment buffer next to the 128 KB table; 70: eight per
# thread, table in global memory; 128: the largest supported cluster (cluster.cu).  MAGPY_B200_CLUSTER_KERNEL=simt keeps
# the clusters the DMMA kernel would take by default on the scalar path.
@pytest.mark.parametrize('N,interactions,renorm', [(2, True, False), (3, True, True), (4, True, True), (5, False, False),
                                                   (7, True, True), (8, True, False), (20, True, False), (40, True, False),
                                                   (64, True, False), (70, True, True), (128, True, False)])
def test_heun_cluster_injected(orc, core, N, interactions, renorm, monkeypatch):
    monkeypatch.setenv('MAGPY_B200_CLUSTER_KERNEL', 'simt')
    rng = np.random.default_rng(N)
    c = ol.make_case(N=N, radius=7e-9 * (1 + 0.1 * rng.random(N)), anisotropy=1e5 * (1 + 0.2 * rng.random(N)),
                     dt=1e-13, t_end=3e-11, S=25, interactions=interactions, renorm=renorm, field_shape='sine',
                     H0=1e4, f=1e10, T=330.0, rng=rng)
    seeds = np.arange(1, 36 if N < 100 else 9) * 101     # 35 members: ragged last CTA (8 for the largest cluster: the oracle is dense)
    t, fl, ref, out, _ = injected_pair(orc, core, c, seeds, per_member=(N <= 4))
    assert out['stats']['kernel'] == ('heun_small' if N <= 7 else 'heun_cluster')
    assert_traj(ref, out, c)


# The same clusters through K2m (cluster_mma.cu): dipolar field as D[3N x 3N] . M[3N x members] on DMMA.8x8x4.
# 8: one particle group, 8 member halves per CTA; 12 / 20: padded last group (forced: the default keeps the scalar
# kernel for badly filled groups); 24, 40: two moment buffers; 56: seven groups; 64: ONE moment buffer next to the 162 KB matrix.
# 140 members: several CTAs, ragged last one, both member halves and all column tiles populated.
# 70, 128: the matrix read from global memory.
@pytest.mark.parametrize('N,renorm,members', [(8, False, 140), (12, True, 35), (20, False, 35), (24, True, 70), (40, False, 35),
                                              (56, False, 20), (64, False, 35), (64, True, 9), (70, True, 20), (128, False, 8)])
def test_heun_cluster_mma_injected(orc, core, N, renorm, members, monkeypatch):
    if N in (12, 20):
        monkeypatch.setenv('MAGPY_B200_CLUSTER_KERNEL', 'mma')
    else:
        monkeypatch.delenv('MAGPY_B200_CLUSTER_KERNEL', raising=False)   # the default choice for these sizes
    rng = np.random.default_rng(N)
    c = ol.make_case(N=N, radius=7e-9 * (1 + 0.1 * rng.random(N)), anisotropy=1e5 * (1 + 0.2 * rng.random(N)),
                     dt=1e-13, t_end=3e-11, S=25, interactions=True, renorm=renorm, field_shape='sine',
                     H0=1e4, f=1e10, T=330.0, rng=rng)
    seeds = np.arange(1, members + 1) * 101
    t, fl, ref, out, _ = injected_pair(orc, core, c, seeds, per_member=(N == 24))
    assert out['stats']['kernel'] == 'heun_cluster_mma'
    assert_traj(ref, out, c)


# Beyond 128 particles: the capacity-free kernel (cluster_big.cu: moments in global memory, increments regenerated for the
# corrector).  The reference takes any cluster size (lib/simulation.cpp:182-200).
@pytest.mark.parametrize('N,renorm,members', [(130, True, 35), (192, False, 5)])
def test_heun_cluster_beyond_128_particles(orc, core, N, renorm, members):
    rng = np.random.default_rng(N)
    c = ol.make_case(N=N, radius=7e-9 * (1 + 0.1 * rng.random(N)), anisotropy=1e5 * (1 + 0.2 * rng.random(N)),
                     dt=1e-13, t_end=1.5e-11, S=13, interactions=True, renorm=renorm, field_shape='sine',
                     H0=1e4, f=1e10, T=330.0, rng=rng)
    seeds = np.arange(1, members + 1) * 101
    t, fl, ref, out, _ = injected_pair(orc, core, c, seeds, per_member=(N == 130))
    assert out['stats']['kernel'] == 'heun_cluster_big'
    assert_traj(ref, out, c)


# The implicit scheme beyond 128 particles (cluster_big.cu, imid_cluster_big_kernel: midpoint iterates in two global
# buffers, a per-member bit says which is current).  The oracle's dense (3N)^3 path limits what can be compared at N > 128
# (474 MB of work arrays and a 390 x 390 dgesv per iteration at N = 130), so the kernel is also forced onto smaller
# clusters (MAGPY_B200_CLUSTER_KERNEL=big), where members that converge after different numbers of iterations share a
# CTA; Heun through the same knob for symmetry.
@pytest.mark.parametrize('N,implicit,interactions,members,steps', [(20, True, True, 40, 12), (33, True, False, 7, 6),
                                                                   (20, False, True, 40, 30), (130, True, True, 2, 3)])
def test_capacity_free_cluster_kernels(orc, core, N, implicit, interactions, members, steps, monkeypatch):
    rng = np.random.default_rng(300 + N)
    c = ol.make_case(N=N, radius=7e-9 * (1 + 0.1 * rng.random(N)), anisotropy=1e5 * (1 + 0.2 * rng.random(N)),
                     dt=1e-12 if implicit else 1e-13, t_end=(1e-12 if implicit else 1e-13) * steps, S=steps // 3 + 1,
                     implicit=implicit, interactions=interactions, renorm=(N == 33), field_shape='sine', H0=1e4, f=1e10,
                     T=330.0, rng=rng)
    seeds = np.arange(1, members + 1) * 29
    if N <= 128:
        monkeypatch.setenv('MAGPY_B200_CLUSTER_KERNEL', 'big')
    t, fl, ref, out, newton = injected_pair(orc, core, c, seeds, per_member=(N == 20))
    assert out['stats']['kernel'] == ('imid_cluster_big' if implicit else 'heun_cluster_big')
    assert_traj(ref, out, c)
    if implicit:
        assert out['stats']['newton_iterations'] == sum(n[0] - n[2] for n in newton)     # identical iteration counts
        assert out['stats']['newton_failures'] == 0
        if N == 20:   # the default kernel for this size on the same case: same iterates
            monkeypatch.delenv('MAGPY_B200_CLUSTER_KERNEL')
            t, fl, ref, out2, newton = injected_pair(orc, core, c, seeds, per_member=True)
            assert out2['stats']['kernel'] == 'imid_cluster_mma'
            assert out2['stats']['newton_iterations'] == out['stats']['newton_iterations']
            assert np.abs(out2['trajectories'] - out['trajectories']).max() / c.Ms < 1e-11


@pytest.mark.parametrize('N,renorm,gauss', [(24, False, 'f32p'), (33, True, 'f64'), (64, False, 'f64')])
def test_mma_and_scalar_cluster_kernels_agree_member_by_member(core, N, renorm, gauss, monkeypatch):
    """Production noise (in-kernel Philox), 300 members, 400 steps, polydisperse: the matrix-product kernel and the
    scalar kernel consume the same increments and differ only in summation order and FMA grouping of the dipolar
    field — per-member trajectories agree far below the parity bar, the fused sums likewise.  (Above 32 particles the
    scalar kernel scales unit fp32 draws in fp64 while K2m folds the amplitude into the fp32 Box-Muller radius — the
    same increments up to fp32 rounding, 6e-8 — so those sizes are compared on the fp64 Gaussian stream.)"""
    rng = np.random.default_rng(7 * N)
    c = ol.make_case(N=N, radius=7e-9 * (1 + 0.1 * rng.random(N)), anisotropy=1e5 * (1 + 0.2 * rng.random(N)),
                     dt=1e-13, t_end=4e-11, S=9, interactions=True, renorm=renorm, field_shape='sine', H0=1e4, f=1e10,
                     T=330.0, rng=rng)
    seeds = rng.integers(1, 2 ** 31 - 1, 300)
    out = {}
    for kern in ('simt', 'mma'):
        monkeypatch.setenv('MAGPY_B200_CLUSTER_KERNEL', kern)
        out[kern] = gpu_run(core, c, seeds, stream_offset=5, gauss=gauss)
    assert out['simt']['stats']['kernel'] == 'heun_cluster' and out['mma']['stats']['kernel'] == 'heun_cluster_mma'
    err = np.abs(out['mma']['trajectories'] - out['simt']['trajectories']).max() / c.Ms
    assert err < 1e-11, err
    assert np.abs(out['mma']['sums'] - out['simt']['sums']).max() / (c.Ms ** 2 * len(seeds) * N * N) < 1e-13


@pytest.mark.parametrize('N', [16, 64])
def test_mma_member_distribution_is_invisible(core, N):
    """K2m gives whole waves of CTAs a full set of members and spreads the rest over all SMs in column tiles of 8:
    a member's trajectory must not depend on where it lands.  A run large enough for full CTAs plus a ragged tail is
    compared, member by member, with small runs of slices of the same ensemble (same seeds, stream_offset)."""
    rng = np.random.default_rng(N)
    c = ol.make_case(N=N, radius=7e-9, anisotropy=1e5, dt=1e-13, t_end=1.2e-12, S=4, interactions=True, T=330.0, rng=rng)
    per_cta = 32 if N == 64 else 128
    R = 148 * 2 * per_cta + 3 * per_cta // 2 + 5 if N == 16 else 148 * per_cta + 45
    seeds = rng.integers(1, 2 ** 31 - 1, R)
    big = gpu_run(core, c, seeds, return_trajectories=False)
    assert big['stats']['kernel'] == 'heun_cluster_mma'
    for lo, hi in ((0, 1), (0, 40), (per_cta - 3, per_cta + 9), (R - 50, R)):   # (0, 1): a single member (core.simulate's case)
        small = gpu_run(core, c, seeds[lo:hi], stream_offset=lo, return_trajectories=False)
        assert np.array_equal(big['final'][lo:hi], small['final'])
    # fused ensemble sums of the big run against its own final states (last sample = final state)
    M = big['final'].sum(axis=1)
    want = np.array([M[:, 0].sum(), M[:, 1].sum(), M[:, 2].sum(), (M[:, 2] ** 2).sum()])
    assert np.allclose(big['sums'][-1], want, rtol=1e-11)


# 8 and more particles: the matrix-product kernel (cluster_mma_imid.cu), 64 = the full shared-memory matrix; the scalar
# kernel (cluster.cu) is kept and tested through MAGPY_B200_CLUSTER_KERNEL=simt (the oracle's dense (3N)^3 path limits
# members and steps at N >= 32)
@pytest.mark.parametrize('N,interactions', [(2, True), (3, False), (4, True), (5, True), (8, True), (12, True), (16, False),
                                            (32, True), (40, True), (64, True)])
def test_implicit_cluster_injected(orc, core, N, interactions, monkeypatch):
    rng = np.random.default_rng(100 + N)
    c = ol.make_case(N=N, radius=7e-9 * (1 + 0.1 * rng.random(N)), anisotropy=1e5 * (1 + 0.2 * rng.random(N)),
                     dt=1e-12, t_end=4e-11, S=21, implicit=True, interactions=interactions, T=330.0, rng=rng)
    seeds = np.arange(1, 34 if N < 32 else 7 if N == 32 else 4) * 13   # the oracle's dense (3N)^3 implicit path is slow at N >= 32
    if N > 32:
        c['t_end'], c['S'] = 1.2e-11 if N < 64 else 6e-12, 7 if N < 64 else 4
    t, fl, ref, out, newton = injected_pair(orc, core, c, seeds, per_member=(N in (2, 4, 12)))
    assert_traj(ref, out, c)
    assert out['stats']['newton_iterations'] == sum(n[0] - n[2] for n in newton)     # identical iteration counts
    assert out['stats']['kernel'] == ('imid_small' if N == 2 else 'imid_warps' if N <= 4 else 'imid_cluster' if N <= 16
                                      else 'imid_cluster_mma')
    if N in (8, 12, 16, 40):   # the other cluster kernel on the same case (default: scalar up to 16 particles, matrix product above)
        other = 'mma' if N <= 16 else 'simt'
        monkeypatch.setenv('MAGPY_B200_CLUSTER_KERNEL', other)
        t, fl, ref, out2, newton = injected_pair(orc, core, c, seeds, per_member=(N == 12))
        assert out2['stats']['kernel'] == ('imid_cluster_mma' if other == 'mma' else 'imid_cluster')
        assert_traj(ref, out2, c)
        assert out2['stats']['newton_iterations'] == out['stats']['newton_iterations']
        monkeypatch.delenv('MAGPY_B200_CLUSTER_KERNEL')
    if N in (2, 3, 4):   # the mappings of small clusters: one thread per cluster, one lane per particle (N = 2, 4), one warp per
        # particle (shared-memory exchange of the iterates; the default for small ensembles of trimers / tetramers)
        for mapping, kernel in (('thread', 'imid_small'), ('split', 'imid_split'), ('warps', 'imid_warps')):
            if mapping == 'split' and N == 3:
                continue
            monkeypatch.setenv('MAGPY_B200_SMALL_KERNEL', mapping)
            t, fl, ref, out2, newton = injected_pair(orc, core, c, seeds, per_member=(N in (2, 4)))
            assert out2['stats']['kernel'] == kernel
            assert_traj(ref, out2, c)
            assert out2['stats']['newton_iterations'] == out['stats']['newton_iterations']
            assert out2['stats']['newton_failures'] == 0


@pytest.mark.parametrize('N,axis_z', [(1, True), (1, False), (2, False), (4, False), (12, False)])
def test_exact_newton_mode_is_the_converged_reference_iteration(orc, core, N, axis_z, monkeypatch):
    """`implicit_newton='exact'` (opt-in) solves the reference's implicit-midpoint equation by Newton's method with the
    exact Jacobian.  The reference's own quasi-Newton iteration converges (linearly) to the same root, so the ORACLE
    run with eps = 1e-14 — the reference algorithm, merely iterated to convergence — is the pathwise comparator: same
    injected Wiener stream, trajectories equal to 1e-11, in 3-4 iterations per step instead of ~35 (and ~20 at the
    default eps = 1e-9, where the reference's truncated iterate sits ~1e-9 from the root)."""
    rng = np.random.default_rng(300 + N)
    c = ol.make_case(N=N, radius=7e-9 * (1 + 0.1 * rng.random(N)), anisotropy=1e5 * (1 + 0.2 * rng.random(N)),
                     dt=1e-12, t_end=6e-11, S=16, implicit=True, interactions=True, T=330.0, field_shape='sine', H0=1e4,
                     f=1e10, eps=1e-14, rng=rng, **(dict(axis=[[0, 0, 1.0]], m0=[[0.6, 0, 0.8]]) if axis_z else {}))
    seeds = np.arange(1, 20) * 17
    n_steps = ol.steps_executed(orc, c)
    dW = np.stack([ol.mt_normal(orc, int(s), n_steps * 3 * N).reshape(n_steps, 3 * N) for s in seeds])
    ref = np.stack([ol.oracle_simulate(orc, c, seed=int(s))[2] for s in seeds])
    fast = gpu_run(core, ol.Case(dict(c, eps=1e-9)), seeds, dW=dW, implicit_newton='exact')
    assert np.abs(fast['trajectories'] - ref).max() / c.Ms < 1e-11
    per_step = fast['stats']['newton_iterations'] / (len(seeds) * fast['stats']['steps_per_member'])
    assert per_step <= 5.0 and fast['stats']['newton_failures'] == 0
    # and the default mode at the default tolerance is the truncated iterate: further from the root, ~20 iterations
    slow = gpu_run(core, ol.Case(dict(c, eps=1e-9)), seeds, dW=dW)
    assert slow['stats']['newton_iterations'] > 4 * fast['stats']['newton_iterations']
    assert 1e-11 < np.abs(slow['trajectories'] - ref).max() / c.Ms < 1e-7
    with pytest.raises(KeyError):
        gpu_run(core, c, seeds, dW=dW, implicit_newton='broyden')
    if N == 4:   # small ensembles of tetramers default to one warp per particle; the other two mappings as well
        assert fast['stats']['kernel'] == 'imid_warps'
        for mapping, kernel in (('thread', 'imid_small'), ('split', 'imid_split')):
            monkeypatch.setenv('MAGPY_B200_SMALL_KERNEL', mapping)
            fast2 = gpu_run(core, ol.Case(dict(c, eps=1e-9)), seeds, dW=dW, implicit_newton='exact')
            assert fast2['stats']['kernel'] == kernel
            assert np.abs(fast2['trajectories'] - ref).max() / c.Ms < 1e-11


@pytest.mark.parametrize('field_shape,axis,renorm,chunk', [('sine', (0, 0, 1.0), True, 7), ('sine', (0.6, 0, 0.8), False, 13),
                                                           ('square', (0.6, 0, 0.8), True, None), ('sine', (0, 0, 1.0), False, 5),
                                                           ('constant', (0, 0, 1.0), False, None)])
def test_heun_single_register_and_latency_variants_are_bit_identical(core, field_shape, axis, renorm, chunk, monkeypatch):
    """heun_single_kernel has three instantiations of its production (packed-noise) form: free register allocation, 7
    resident CTAs per SM (shards that fit one such wave), and the latency variant for small ensembles (applied-field table
    entries fetched one step pair ahead).  Same Philox counters, same arithmetic: bit-identical trajectories, sums and
    final states, for ragged ensembles, odd launch boundaries, samples finer than the step."""
    c = ol.make_case(N=1, dt=1e-13, t_end=4.37e-11, S=57, field_shape=field_shape, H0=2e4, f=4e10, renorm=renorm,
                     axis=[list(axis)], m0=[[1.0, 0, 0]])
    if chunk:
        monkeypatch.setenv('MAGPY_B200_MAX_CHUNK_STEPS', str(chunk))
    for R, S in ((1000, 57), (37, 57), (64, 600)):          # S = 600 > steps: repeated samples (zero-order hold)
        cc = ol.Case(dict(c, S=S))
        seeds = np.arange(R) * 7 + 1
        outs = {}
        for v in ('1', '7', '100'):
            monkeypatch.setenv('MAGPY_B200_K1_MIN_BLOCKS', v)
            outs[v] = gpu_run(core, cc, seeds, stream_offset=5)
            assert outs[v]['stats']['kernel'] == 'heun_single' and outs[v]['stats']['kernel_variant'] == int(v)
        for v in ('7', '100'):
            assert np.array_equal(outs['1']['trajectories'], outs[v]['trajectories'])
            assert np.array_equal(outs['1']['final'], outs[v]['final'])
            assert np.array_equal(outs['1']['sums'], outs[v]['sums'])
        # K1s (heun_single_split.cu): an integrator warp fed by a generator warp through a shared-memory ring — the same
        # arithmetic per member (bit-identical trajectories and final states); the ensemble sums are formed per 32 members
        # instead of per 128, i.e. in another fixed order
        monkeypatch.delenv('MAGPY_B200_K1_MIN_BLOCKS')
        monkeypatch.setenv('MAGPY_B200_K1_SPLIT', '1')
        sp = gpu_run(core, cc, seeds, stream_offset=5)
        monkeypatch.delenv('MAGPY_B200_K1_SPLIT')
        assert sp['stats']['kernel'] == 'heun_single' and sp['stats']['kernel_variant'] == 300
        assert np.array_equal(outs['1']['trajectories'], sp['trajectories'])
        assert np.array_equal(outs['1']['final'], sp['final'])
        assert np.allclose(outs['1']['sums'], sp['sums'], rtol=1e-13, atol=1e-13 * np.abs(outs['1']['sums']).max())
    small = 100 if field_shape != 'constant' else 1          # beyond K1s: the latency variant when there is a field table
    assert gpu_run(core, c, np.arange(100))['stats']['kernel_variant'] == 300         # K1s: the default up to 64 members per SM
    assert gpu_run(core, c, np.arange(12000))['stats']['kernel_variant'] == small


@pytest.mark.parametrize('field_shape,axis,renorm,chunk', [('constant', (0, 0, 1.0), False, None), ('sine', (0, 0, 1.0), True, 4999),
                                                           ('sine', (0.6, 0, 0.8), False, None), ('square', (0.6, 0, 0.8), True, 7000)])
def test_balanced_persistent_heun_kernel_is_bit_identical(core, field_shape, axis, renorm, chunk, monkeypatch):
    """Multi-wave single-particle Heun shards run as a persistent kernel over (time segment, block of 128 members) tasks
    (heun_single_balanced.cu: state parked in HBM between segments, release / acquire flag per block).  Same Philox
    counters, same arithmetic: bit-identical trajectories, sums and final states to heun_single_kernel — forced here on
    small ensembles (one CTA juggling many tasks, ragged last block), with samples finer and coarser than a segment and
    odd launch boundaries; the default switches it on above one wave of resident CTAs."""
    c = ol.make_case(N=1, dt=1e-13, t_end=1.2e-9, S=57, field_shape=field_shape, H0=2e4, f=4e9, renorm=renorm,
                     axis=[list(axis)], m0=[[1.0, 0, 0]])                      # 12,000 steps: 11 segments of >= 1024 steps
    if chunk:
        monkeypatch.setenv('MAGPY_B200_MAX_CHUNK_STEPS', str(chunk))
    for R, S in ((1000, 57), (37, 3), (300, 20000)):          # S = 20000 > steps: repeated samples (zero-order hold)
        cc = ol.Case(dict(c, S=S))
        seeds = np.arange(R) * 7 + 1
        monkeypatch.setenv('MAGPY_B200_K1_MIN_BLOCKS', '1')
        monkeypatch.setenv('MAGPY_B200_K1_BALANCE', '0')
        a = gpu_run(core, cc, seeds, stream_offset=5)
        monkeypatch.setenv('MAGPY_B200_K1_BALANCE', '1')
        b = gpu_run(core, cc, seeds, stream_offset=5)
        assert a['stats']['kernel_variant'] == 1 and b['stats']['kernel_variant'] == 200
        assert np.array_equal(a['trajectories'], b['trajectories'])
        assert np.array_equal(a['final'], b['final'])
        assert np.array_equal(a['sums'], b['sums'])
    for k in ('MAGPY_B200_K1_MIN_BLOCKS', 'MAGPY_B200_K1_BALANCE', 'MAGPY_B200_MAX_CHUNK_STEPS'):
        monkeypatch.delenv(k, raising=False)
    big = gpu_run(core, ol.Case(dict(c, t_end=3e-10, S=5)), np.arange(200000), return_trajectories=False)
    assert big['stats']['kernel_variant'] == 200                                # 1563 blocks > 888 resident CTAs
    monkeypatch.setenv('MAGPY_B200_K1_BALANCE', '0')
    ref = gpu_run(core, ol.Case(dict(c, t_end=3e-10, S=5)), np.arange(200000), return_trajectories=False)
    assert ref['stats']['kernel_variant'] == 1 and np.array_equal(ref['final'], big['final']) and np.array_equal(ref['sums'], big['sums'])


def test_adjugate_solve_against_pivoted_elimination(core):
    """The implicit kernels solve each particle's 3x3 quasi-Newton system by the adjugate (llg_math.cuh) where the
    reference calls dgesv (pivoted elimination, lib/optimisation.cpp:134).  On the matrices of the iteration (I + O(dt))
    no pivoting happens and the two agree to a few ulp; this checks the regime where the pivot order WOULD matter:
    random matrices with condition numbers up to 1e8, matrices with a zero or tiny leading entry (elimination needs a
    row swap, the adjugate needs nothing), and exactly singular ones (dgesv info > 0 <-> ok = False).  The distance
    to LAPACK's solution stays within a modest multiple of cond * eps — the scale of the forward error pivoted elimination
    itself has (observed maximum over these 4096 systems: 69 cond eps; bar 256)."""
    rng = np.random.default_rng(42)
    n = 4096
    U, _ = np.linalg.qr(rng.normal(size=(n, 3, 3)))
    V, _ = np.linalg.qr(rng.normal(size=(n, 3, 3)))
    sv = np.stack([np.ones(n), 10.0 ** rng.uniform(-4, 0, n), 10.0 ** rng.uniform(-8, 0, n)], axis=1)
    A = U @ (sv[:, :, None] * np.swapaxes(V, 1, 2))
    A[:512, 0, 0] = 0.0                          # zero pivot in the natural order
    A[512:1024, 0, 0] *= 1e-14                   # tiny pivot
    near_id = np.eye(3) + 0.05 * rng.normal(size=(n, 3, 3))
    A[1024:2048] = near_id[1024:2048]            # what the iteration actually produces
    b = rng.normal(size=(n, 3))
    x, ok = core.solve3(A, b)
    want = np.linalg.solve(A, b[:, :, None])[:, :, 0]          # LAPACK gesv: partial pivoting
    cond = np.linalg.cond(A)
    rel = np.linalg.norm(x - want, axis=1) / np.linalg.norm(want, axis=1)
    assert ok.all()
    assert (rel <= 256 * cond * np.finfo(float).eps).all(), (rel / (cond * np.finfo(float).eps)).max()
    assert rel[1024:2048].max() < 1e-14           # the iteration's own matrices: a few ulp
    S = np.zeros((3, 3, 3)); S[0] = [[1, 2, 3], [2, 4, 6], [0, 1, 5]]; S[1] = 0; S[2] = [[1, 0, 0], [0, 1, 0], [1, 1, 0]]
    _, ok_s = core.solve3(S, np.ones((3, 3)))
    assert not ok_s.any()


@pytest.mark.parametrize('implicit', [False, True])
def test_per_member_radii_single_particle(orc, core, implicit):
    """A size distribution in ONE launch: single-particle members that differ in radius (radius of shape (R, 1) ->
    `radius_stride`), each against the oracle run with its own radius and the reference's Wiener stream of its seed.
    Through EnsembleModel the same ensemble (radius=[...] per member) takes one device call instead of one per size."""
    rng = np.random.default_rng(9)
    R = 40
    radii = rng.uniform(5e-9, 9e-9, R)
    seeds = np.arange(1, R + 1) * 37
    base = ol.make_case(N=1, dt=1e-13 if not implicit else 1e-12, t_end=2e-11 if not implicit else 6e-11, S=17, implicit=implicit,
                        field_shape='sine', H0=1.5e4, f=5e9, axis=[[0.6, 0, 0.8]], m0=[[0, 0, 1.0]])
    n_steps = ol.steps_executed(orc, base)          # the schedule does not depend on the radius
    dW = np.stack([ol.mt_normal(orc, int(s), n_steps * 3).reshape(n_steps, 3) for s in seeds])
    ref = np.stack([ol.oracle_simulate(orc, ol.Case(dict(base, radius=np.array([radii[i]]))), seed=int(seeds[i]))[2]
                    for i in range(R)])
    out = core.simulate_ensemble(radii.reshape(R, 1), base.anisotropy, base.axis, base.m0, base.location, base.Ms, base.alpha,
                                 base.T, False, True, implicit, base.dt, base.t_end, base.S, seeds, field_shape='sine',
                                 field_amplitude=base.H0, field_frequency=base.f, injected_dw=dW)
    assert np.abs(out['trajectories'] - ref).max() / base.Ms <= TOL
    assert out['stats']['kernel_launches'] < 12            # one integration launch, not one per radius
    # the noise amplitude really is per member: the smallest particle wanders furthest from its start
    spread = np.abs(out['trajectories'][:, 0, :, -1] / base.Ms - np.array([0, 0, 1.0])).max(axis=1)
    assert spread[np.argmin(radii)] > spread[np.argmax(radii)]
    with pytest.raises(ValueError):                          # clusters: per-member radii change the geometry-dependent tables
        c2 = ol.make_case(N=2)
        core.simulate_ensemble(np.full((3, 2), 7e-9), c2.anisotropy, c2.axis, c2.m0, c2.location, c2.Ms, c2.alpha, c2.T,
                               False, True, False, c2.dt, c2.t_end, c2.S, [1, 2, 3])
    import magpy_b200 as mp
    model = mp.Model(np.array([7e-9]), base.anisotropy, base.axis, base.m0, base.location, base.Ms, base.alpha, base.T,
                     field_shape='sine', field_frequency=base.f, field_amplitude=base.H0)
    ens = mp.EnsembleModel(R, model, radius=[np.array([r]) for r in radii])
    res = ens.simulate(base.t_end, base.dt, base.S, 1001, implicit_solve=implicit)
    assert len(res.stats) == 1 and res.stats[0]['particle_steps'] == R * n_steps - R   # the product skips the last, never-sampled step
    # per-member temperatures ride on the same per-member sigma: a temperature sweep against the oracle, member by member
    temps = rng.uniform(100.0, 400.0, R)
    ref_t = np.stack([ol.oracle_simulate(orc, ol.Case(dict(base, T=float(temps[i]))), seed=int(seeds[i]))[2] for i in range(R)])
    out_t = core.simulate_ensemble(base.radius, base.anisotropy, base.axis, base.m0, base.location, base.Ms, base.alpha,
                                   temps, False, True, implicit, base.dt, base.t_end, base.S, seeds, field_shape='sine',
                                   field_amplitude=base.H0, field_frequency=base.f, injected_dw=dW)
    assert np.abs(out_t['trajectories'] - ref_t).max() / base.Ms <= TOL
    ens_t = mp.EnsembleModel(R, model, temperature=list(temps), radius=[np.array([r]) for r in radii])
    assert len(ens_t.simulate(base.t_end, base.dt, base.S, 1001, implicit_solve=implicit).stats) == 1


@pytest.mark.parametrize('implicit', [False, True])
@pytest.mark.parametrize('shape', ['sine', 'constant'])
def test_per_member_material_parameters_single_particle(orc, core, implicit, shape):
    """Anisotropy, damping, field amplitude (next to radius and temperature) per member in ONE launch — the
    reference's `EnsembleModel(N, base, anisotropy=[...], damping=[...], ...)` (magpy/model.py:146-156).  Anisotropy
    and damping change the member's reduced time scale, so every thread runs its own dt, noise amplitude and
    zero-order-hold schedule; each member is checked against the oracle run with its own parameters and the reference's
    Wiener stream of its seed.  t_end / dt / (S - 1) is an integer here: every sampling time is an exact tie of the
    schedule, where the members' step counts differ by rounding alone (the reference's per-member outcome is the bar)."""
    rng = np.random.default_rng(11)
    R = 48
    K = rng.uniform(2e4, 9e4, R)
    al = rng.uniform(0.05, 0.5, R)
    H0 = rng.uniform(0.0, 3e4, R)
    rad = rng.uniform(5e-9, 9e-9, R)
    T = rng.uniform(150.0, 400.0, R)
    seeds = np.arange(1, R + 1) * 41
    base = ol.make_case(N=1, dt=1e-13 if not implicit else 1e-12, t_end=2.4e-11 if not implicit else 1.12e-10, S=17,
                        implicit=implicit, field_shape=shape, H0=1.5e4, f=5e9, axis=[[0.6, 0, 0.8]], m0=[[0, 0, 1.0]])
    cases = [ol.Case(dict(base, radius=np.array([rad[i]]), anisotropy=np.array([K[i]]), alpha=float(al[i]), H0=float(H0[i]),
                          T=float(T[i]))) for i in range(R)]
    steps = np.array([ol.steps_executed(orc, c) for c in cases])
    assert len(set(steps.tolist())) > 1, 'the members should not all share one schedule (exact ties)'
    n_max = int(steps.max()) + 2
    dW = np.stack([ol.mt_normal(orc, int(s), n_max * 3).reshape(n_max, 3) for s in seeds])
    ref, newton = [], []
    for c, s in zip(cases, seeds):
        t, fl, m, it, fails = ol.oracle_simulate(orc, c, seed=int(s))
        ref.append(m); newton.append(it)
    ref = np.stack(ref)
    out = core.simulate_ensemble(rad.reshape(R, 1), K.reshape(R, 1), base.axis, base.m0, base.location, base.Ms, al, T,
                                 False, True, implicit, base.dt, base.t_end, base.S, seeds, field_shape=shape,
                                 field_amplitude=H0, field_frequency=base.f, injected_dw=dW)
    err = np.abs(out['trajectories'] - ref).max(axis=(1, 2, 3)) / base.Ms
    assert err.max() <= TOL, (err.argmax(), err.max())
    assert out['stats']['kernel_launches'] < 12
    assert np.array_equal(out['final'], out['trajectories'][..., -1])      # no member steps past its own last sample
    if implicit:   # identical iteration counts (the product skips each member's last, never-sampled step)
        assert out['stats']['newton_iterations'] == sum(n[0] - n[2] for n in newton)
    # the field every member saw: unit waveform times its amplitude
    if shape == 'sine':
        assert np.allclose(out['field'] * H0[3], ol.oracle_simulate(orc, cases[3], seed=1)[1], rtol=1e-12, atol=1e-9)
    # each parameter alone also rides on the per-member arrays
    for kw in (dict(anisotropy=K[:8].reshape(8, 1)), dict(damping=al[:8]), dict(field_amplitude=H0[:8])):
        a = dict(anisotropy=base.anisotropy, damping=base.alpha, field_amplitude=base.H0)
        a.update(kw)
        one = core.simulate_ensemble(base.radius, a['anisotropy'], base.axis, base.m0, base.location, base.Ms, a['damping'],
                                     base.T, False, True, implicit, base.dt, base.t_end, base.S, seeds[:8], field_shape=shape,
                                     field_amplitude=a['field_amplitude'], field_frequency=base.f, injected_dw=dW[:8])
        i = 5
        ci = ol.Case(dict(base, anisotropy=np.array([K[i]]) if 'anisotropy' in kw else base.anisotropy,
                          alpha=float(al[i]) if 'damping' in kw else base.alpha,
                          H0=float(H0[i]) if 'field_amplitude' in kw else base.H0))
        want = ol.oracle_simulate(orc, ci, seed=int(seeds[i]))[2]
        assert np.abs(one['trajectories'][i] - want).max() / base.Ms <= TOL, list(kw)
    # through the public API: one device call for the whole distribution, Philox noise
    import magpy_b200 as mp
    model = mp.Model(np.array([7e-9]), base.anisotropy, base.axis, base.m0, base.location, base.Ms, base.alpha, base.T,
                     field_shape=shape, field_frequency=base.f, field_amplitude=base.H0)
    ens = mp.EnsembleModel(R, model, anisotropy=[np.array([k]) for k in K], damping=list(al), field_amplitude=list(H0),
                           radius=[np.array([r]) for r in rad])
    res = ens.simulate(base.t_end, base.dt, base.S, 1001, implicit_solve=implicit)
    assert len(res.stats) == 1
    assert np.allclose(res.results[7].field, res.results[0].field * (H0[7] / H0[0]) if shape == 'sine' else H0[7])
    with pytest.raises(ValueError):     # the square wave's switching instants depend on the member's time scale
        core.simulate_ensemble(base.radius, K.reshape(R, 1), base.axis, base.m0, base.location, base.Ms, base.alpha, base.T,
                               False, True, implicit, base.dt, base.t_end, base.S, seeds, field_shape='square',
                               field_amplitude=1e4, field_frequency=base.f)


@pytest.mark.parametrize('implicit', [False, True])
def test_per_member_parameters_do_not_depend_on_chunking(core, implicit, monkeypatch):
    """The per-member-parameter kernels carry every member's own step index between launches (the members' schedules differ
    by a step at exact ties): many short launches — sampling launches and pure-advance launches — give bit-identical
    results to one launch, with the Philox stream and with more samples than steps."""
    rng = np.random.default_rng(5)
    R = 200
    K = rng.uniform(2e4, 9e4, R).reshape(R, 1)
    al = rng.uniform(0.05, 0.5, R)
    H0 = rng.uniform(0.0, 3e4, R)
    seeds = np.arange(R) * 3 + 11
    for S, t_end in ((17, 2.4e-11 if not implicit else 1.12e-10), (3, 2.4e-11 if not implicit else 1.12e-10), (500, 3e-11 if not implicit else 1.2e-10)):
        c = ol.make_case(N=1, dt=1e-13 if not implicit else 1e-12, t_end=t_end, S=S, implicit=implicit, field_shape='sine',
                         H0=1.5e4, f=5e9, axis=[[0.6, 0, 0.8]], m0=[[0, 0, 1.0]])

        def run():
            return core.simulate_ensemble(c.radius, K, c.axis, c.m0, c.location, c.Ms, al, c.T, False, True, implicit, c.dt,
                                          c.t_end, c.S, seeds, field_shape='sine', field_amplitude=H0, field_frequency=c.f,
                                          stream_offset=9)
        monkeypatch.delenv('MAGPY_B200_MAX_CHUNK_STEPS', raising=False)
        whole = run()
        for chunk in ('7', '40'):
            monkeypatch.setenv('MAGPY_B200_MAX_CHUNK_STEPS', chunk)
            parts = run()
            assert parts['stats']['kernel_launches'] > whole['stats']['kernel_launches'] + 2
            assert np.array_equal(whole['trajectories'], parts['trajectories']), (S, chunk)
            assert np.array_equal(whole['final'], parts['final'])
            assert np.allclose(whole['sums'], parts['sums'], rtol=1e-13, atol=0)
            if implicit:
                assert whole['stats']['newton_iterations'] == parts['stats']['newton_iterations']
        assert np.array_equal(whole['final'], whole['trajectories'][..., -1])


def test_single_simulate_api_and_schedule_edges(orc, core):
    """core.simulate keeps the reference's dict; sampling finer than the time step repeats states."""
    c = ol.make_case(N=2, dt=1e-12, t_end=1e-11, S=40, implicit=False)    # Ts < dt: zero-order hold repeats
    res = core.simulate(c.radius, c.anisotropy, c.axis, c.m0, c.location, c.Ms, c.alpha, c.T, False, True, False,
                        c.dt, c.t_end, c.S, 1234)
    assert set(res) == {'N', 'time', 'field', 'x', 'y', 'z'} and res['N'] == 2
    t, fl, m, _, _ = ol.oracle_simulate(orc, c, seed=1)
    assert np.array_equal(res['time'], t)
    # identical hold pattern: consecutive samples equal exactly where the oracle's are
    same_ref = np.all(m[:, :, 1:] == m[:, :, :-1], axis=(0, 1))
    got = np.stack([[res['x'][i], res['y'][i], res['z'][i]] for i in range(2)])
    same_got = np.all(got[:, :, 1:] == got[:, :, :-1], axis=(0, 1))
    assert np.array_equal(same_ref, same_got)
    assert np.allclose(got[:, :, 0], c.m0 * c.Ms)
    with pytest.raises(KeyError):
        core.simulate(c.radius, c.anisotropy, c.axis, c.m0, c.location, c.Ms, c.alpha, c.T, False, True, False,
                      c.dt, c.t_end, c.S, 1234, field_shape='triangle')
    with pytest.raises(ValueError):
        core.simulate(c.radius, c.anisotropy, c.axis, c.m0, c.location, c.Ms, c.alpha, c.T, False, True, False,
                      c.dt, c.t_end, 1, 1234)


@pytest.mark.parametrize('implicit', [False, True])
def test_coarsened_noise_is_the_sum_of_the_fine_stream(core, implicit):
    """noise_coarsen_log2 = L: step s is driven by 2^(-L/2) times the sum of the packed Philox increments of the
    fine steps s 2^L ... (s+1) 2^L - 1 (test/convergence/task5.cpp:150-158) — checked against an injected-noise
    run fed with exactly those sums, built on the host from the device's own fine stream."""
    L, n_coarse, R = 3, 40, 6
    n_fine = n_coarse << L
    c = ol.make_case(N=1, radius=6e-9, dt=2e-13, t_end=2e-13 * n_coarse * (1 + 1e-9), S=2, implicit=implicit,
                     axis=[[0, 0, 1.0]], m0=[[0.6, 0, 0.8]])
    seeds = np.array([11, 22, 33, 44, 55, 66])
    offset = 1000
    fine = np.stack([core.gaussians(int(s), offset + i, 0, 1, n_fine, gauss='f32p') for i, s in enumerate(seeds)])
    coarse = fine.reshape(R, n_coarse, 1 << L, 3).sum(axis=2) / np.sqrt(1 << L)
    coarse = np.concatenate([coarse, np.zeros((R, 2, 3))], axis=1)
    a = gpu_run(core, c, seeds, dW=coarse)
    b = gpu_run(core, c, seeds, stream_offset=offset, noise_coarsen_log2=L)
    assert a['stats']['steps_per_member'] == b['stats']['steps_per_member'] == n_coarse
    assert np.abs(a['final'] - b['final']).max() / c.Ms < 1e-12
    # L = 0 is the production stream itself, up to fp32 rounding of every increment: the production kernels fold the
    # amplitude sigma sqrt(dt) into the fp32 Box-Muller radius, the coarsened sums scale unit draws in fp64
    c0 = ol.Case(dict(c, dt=c.dt / (1 << L)))
    f = np.concatenate([fine, np.zeros((R, 2, 3))], axis=1)
    a0 = gpu_run(core, c0, seeds, dW=f)
    b0 = gpu_run(core, c0, seeds, stream_offset=offset)
    assert np.abs(a0['final'] - b0['final']).max() / c.Ms < 1e-7
    with pytest.raises(ValueError):
        gpu_run(core, ol.make_case(N=2), seeds, noise_coarsen_log2=1)


@pytest.mark.parametrize('N,implicit', [(1, False), (3, False), (6, False), (12, False), (40, False), (1, True), (2, True), (6, True),
                                        (12, True), (24, True), (130, False)])
def test_every_kernel_family_consumes_the_same_philox_stream(core, N, implicit, monkeypatch):
    """The increment of (seed, member, particle, step) is one function — the one `core.gaussians` exposes and
    tests/test_parity_gpu.py checks against the oracle — whichever kernel consumes it: pipelined pairs (single
    particle), blocks carried across sample boundaries (one thread per cluster), pairwise or per-step access
    (shared-memory clusters), with odd chunk and sample boundaries.  A production (Philox) run therefore equals an
    injected-noise run fed with `core.gaussians` up to the fp32 rounding of the Heun kernels' folded amplitude."""
    monkeypatch.setenv('MAGPY_B200_MAX_CHUNK_STEPS', '7')
    rng = np.random.default_rng(50 + N)
    c = ol.make_case(N=N, radius=7e-9, anisotropy=1e5, dt=2e-13 if not implicit else 1e-12, S=6, implicit=implicit,
                     T=330.0, rng=rng)
    n_steps = 45
    c['t_end'] = c.dt * n_steps * (1 + 1e-9)
    seeds = np.array([3, 1 << 33, 77, 2 ** 31 - 1, 12345])
    offset = 17
    dW = np.zeros((len(seeds), n_steps + 2, 3 * N))
    for i, s in enumerate(seeds):
        for p in range(N):
            dW[i, :n_steps, 3 * p:3 * p + 3] = core.gaussians(int(s), offset + i, p, 1, n_steps, gauss='f32p')
    a = gpu_run(core, c, seeds, dW=dW)
    b = gpu_run(core, c, seeds, stream_offset=offset)
    assert a['stats']['steps_per_member'] == b['stats']['steps_per_member'] == n_steps
    assert b['stats']['kernel_launches'] > 6
    err = np.abs(a['trajectories'] - b['trajectories']).max() / c.Ms
    assert err < (1e-7 if not implicit else 1e-11), err
