"""``Model`` / ``EnsembleModel`` with the interface of the reference's ``magpy/model.py``.

``Model`` is unchanged in meaning (magpy/model.py:16-115).  ``EnsembleModel.simulate``
(magpy/model.py:159-208) keeps its signature and the derivation of the per-member seeds
from `random_state`, but replaces the joblib process pool by ONE batched device call per
group of members that share geometry and material (members may differ freely in
`anisotropy_axis` and `magnetisation_direction` inside a group).
"""
import numpy as np

from . import core
from .results import Results, EnsembleResults

_PER_MEMBER_FAST = ('anisotropy_axis', 'magnetisation_direction')
_MAX_CONCURRENT_PLANS = 32   # parameter groups in flight at once (each holds its own device buffers)
_TRAJ_BYTES_AUTO = 2 << 30


class Model:
    """A cluster of interacting magnetic nanoparticles (magpy/model.py:16-68).

    Args:
        radius (list of double): radius of each spherical particle [m]
        anisotropy (list of double): uniaxial anisotropy constant of each particle [J/m^3]
        anisotropy_axis (list of ndarray[double,3]): unit anisotropy axis of each particle
        magnetisation_direction (list of ndarray[double,3]): initial magnetisation direction
        location (list of ndarray[double,3]): particle coordinates [m]
        magnetisation (double): saturation magnetisation of all particles [A/m]
        damping (double): damping constant of all particles
        temperature (double): ambient temperature [K]
        field_shape (str, optional): 'constant', 'square' or 'sine' (applied along z)
        field_frequency (double, optional): [Hz]
        field_amplitude (double, optional): [A/m]
    """
    def __init__(self, radius, anisotropy, anisotropy_axis, magnetisation_direction, location,
                 magnetisation, damping, temperature, field_shape='constant', field_frequency=0.0,
                 field_amplitude=0.0):
        self.radius = np.array(radius)
        self.anisotropy = np.array(anisotropy)
        self.anisotropy_axis = np.array(anisotropy_axis)
        self.magnetisation_direction = np.array(magnetisation_direction)
        self.location = np.array(location)
        self.magnetisation = magnetisation
        self.damping = damping
        self.temperature = temperature
        self.field_shape = field_shape
        self.field_frequency = field_frequency
        self.field_amplitude = field_amplitude

    def simulate(self, end_time, time_step, max_samples, seed=1001, renorm=False, interactions=True,
                 implicit_solve=True, implicit_tol=1e-9):
        """Simulate the cluster; same arguments and defaults as magpy/model.py:70-115."""
        res = core.simulate(
            np.ascontiguousarray(self.radius, dtype=np.float64).reshape(-1),
            np.ascontiguousarray(self.anisotropy, dtype=np.float64).reshape(-1),
            np.ascontiguousarray(self.anisotropy_axis, dtype=np.float64).reshape(-1, 3),
            np.ascontiguousarray(self.magnetisation_direction, dtype=np.float64).reshape(-1, 3),
            np.ascontiguousarray(self.location, dtype=np.float64).reshape(-1, 3),
            self.magnetisation, self.damping, self.temperature, renorm, interactions,
            implicit_solve, time_step, end_time, max_samples, seed,
            self.field_shape, self.field_amplitude, self.field_frequency, implicit_tol)
        return Results(**res)


def _group_key(params, keys):
    parts = []
    for k in keys:
        v = params[k]
        if isinstance(v, np.ndarray):
            parts.append((k, v.shape, v.tobytes()))
        else:
            parts.append((k, v))
    return tuple(parts)


class EnsembleModel:
    """Ensemble of particle clusters (magpy/model.py:118-156).

    Args:
        N (int): number of clusters in the ensemble
        base_model (Model): default parameters of every member
        **kwargs: a Model parameter name and a list of `N` values, one per member
    """
    def __init__(self, N, base_model, **kwargs):
        self.ensemble_size = N
        self.base_model_params = base_model.__dict__
        for key, vals in kwargs.items():
            if key not in self.base_model_params:
                raise TypeError("__init__() got an unexpected keyword argument '%s'" % key)
            if len(vals) < N:
                raise IndexError('list index out of range')
        self._overrides = kwargs
        self._models = None
        self._seed_cache = {}

    @property
    def model_params(self):
        return [dict(self.base_model_params, **{key: self._overrides[key][i] for key in self._overrides})
                for i in range(self.ensemble_size)]

    @property
    def models(self):
        # built on demand: the reference deep-copies N Models eagerly (magpy/model.py:149-156)
        if self._models is None:
            self._models = [Model(**params) for params in self.model_params]
        return self._models

    def _member_seeds(self, random_state):
        # magpy/model.py:202-203.  The draw of 1M seeds takes ~5 ms — 7 % of a pass on eight GPUs — so repeated calls
        # with the same random_state reuse it, leaving numpy's global generator in the state the draw would have left.
        key = (random_state, self.ensemble_size)
        hit = self._seed_cache.get(key) if isinstance(random_state, (int, np.integer)) else None
        if hit is None:
            np.random.seed(random_state)
            seeds = np.random.randint(np.iinfo(np.int32).max, size=self.ensemble_size)
            if isinstance(random_state, (int, np.integer)):
                self._seed_cache.clear()
                self._seed_cache[key] = (seeds, np.random.get_state())
            return seeds
        np.random.set_state(hit[1])
        return hit[0]

    def simulate(self, end_time, time_step, max_samples, random_state, renorm=False, interactions=True,
                 n_jobs=1, implicit_solve=True, implicit_tol=1e-9, device=0, stream_offset=0,
                 return_trajectories=None, gauss='f32p', shard=None, devices=None, implicit_newton='reference',
                 comm=None):
        """Simulate every member; arguments up to `implicit_tol` as magpy/model.py:159-208.

        `n_jobs` is accepted and ignored (the ensemble runs as one device launch).
        Extra keyword arguments (defaults keep the reference behaviour):
            device (int): CUDA device ordinal.
            stream_offset (int): global index of member 0 (used when an ensemble is sharded).
            return_trajectories (bool|None): keep per-member trajectories; None = keep them
                when they take less than 2 GiB.
            gauss ('f32p'|'f32'|'f64'): Gaussian transform of the in-kernel Philox stream (packed fp32
                Box-Muller, one Philox block per two steps; fp32 with 32-bit uniforms; fp64).
            devices (list of int|'all'|None): shard the members over several GPUs of this box from this one
                process (one stream per device, ensemble sums added on the host); overrides `device`.
            implicit_newton ('reference'|'exact'): implicit midpoint only.  'reference' reproduces the reference's
                quasi-Newton iteration iterate by iterate (~20 iterations per step); 'exact' is Newton's method with the
                exact Jacobian of the midpoint residual (~3 iterations, several times faster): the same scheme solved to
                a tighter residual, so paths differ from the reference's at the 1e-9 level per step.
            shard ((rank, world_size)|None): integrate only this rank's contiguous slice of the
                members (magpy_b200.sharding.shard_bounds) and all-reduce the ensemble sums over `comm`;
                per-member outputs then cover the local slice.
            comm (core.Comm|None): the communicator of a sharded run (one process per GPU).  None = the one built
                from the launcher's environment (`sharding.world_comm`, torchrun's RANK / WORLD_SIZE / MASTER_*);
                a sharded run over more than one rank without a communicator raises.  With a `core.Comm` and a
                single parameter group the all-reduce runs on the device inside the library call (one
                ncclAllReduce); any other object needs an `allreduce_sums(ndarray)` method and is applied to the
                host sums.
        """
        R = self.ensemble_size
        seeds = self._member_seeds(random_state)
        lo, hi = 0, R
        if shard is not None:
            from .sharding import shard_bounds
            lo, hi = shard_bounds(R, shard[1], shard[0])
            if hi == lo:
                raise ValueError('ensemble of %d members cannot be sharded over %d ranks' % (R, shard[1]))
            from .sharding import resolve_comm
            comm = resolve_comm(shard, comm, device)
        elif comm is not None:
            raise ValueError('comm= needs shard=(rank, world_size)')
        device_comm = comm if isinstance(comm, core.Comm) else None
        base = self.base_model_params
        N = int(np.asarray(base['radius']).reshape(-1).shape[0])
        if return_trajectories is None:
            return_trajectories = (hi - lo) * N * 3 * int(max_samples) * 8 <= _TRAJ_BYTES_AUTO

        # per-member radii / temperatures / anisotropy constants / damping / field amplitudes of single-particle members go
        # to the device as arrays too (one launch for a size or anisotropy distribution, a temperature or amplitude sweep).
        # Anisotropy and damping change the member's time scale; the library evaluates each member's own schedule but not
        # a square wave's switching instants per member, so with a square field they fall back to one group per value.
        fast = _PER_MEMBER_FAST
        if N == 1:
            fast += ('radius', 'temperature')
            if gauss == 'f32p':      # the per-member-parameter kernels exist for the production noise mode
                fast += ('field_amplitude',)
                if base['field_shape'] != 'square' and 'field_shape' not in self._overrides:
                    fast += ('anisotropy', 'damping')
        other_keys = [k for k in self._overrides if k not in fast]
        if other_keys:
            groups = {}
            for i in range(lo, hi):
                params = {k: np.asarray(self._overrides[k][i]) if isinstance(base[k], np.ndarray)
                          else self._overrides[k][i] for k in other_keys}
                groups.setdefault(_group_key(params, other_keys), []).append(i)
            groups = [np.asarray(v) for v in groups.values()]
        else:
            groups = [np.arange(lo, hi)]

        S = int(max_samples)
        single = len(groups) == 1     # the usual case: one device call, its output arrays are the result
        traj = final = None
        if not single:
            traj = np.empty((hi - lo, N, 3, S)) if return_trajectories else None
            final = np.empty((hi - lo, N, 3))
        sums = np.zeros((S, 4))
        state = dict(time=None, field=None, traj=traj, final=final, sums=sums)
        group_fields, group_of = [], np.zeros(hi - lo, dtype=np.int64)   # the applied field each member saw
        member_amp = None
        if N == 1 and 'field_amplitude' in self._overrides and 'field_amplitude' in fast:
            # the library returns the waveform for 1 A/m; member i saw field_amplitude[i] times it
            member_amp = np.asarray([self._overrides['field_amplitude'][i] for i in range(lo, hi)], dtype=np.float64)
        pending = []

        def consume(out, idx):
            if state['time'] is None:
                state['time'], state['field'] = out['time'], out['field']
                if member_amp is not None:   # the ensemble-level field is member 0's, as the reference's results[0].field
                    state['field'] = out['field'] * member_amp[idx[0] - lo]
            if single:
                state['traj'], state['final'], state['sums'] = out['trajectories'], out['final'], out['sums']
                group_fields.append(out['field'])
            else:
                if return_trajectories:
                    state['traj'][idx - lo] = out['trajectories']
                state['final'][idx - lo] = out['final']
                state['sums'] += out['sums']
                group_of[idx - lo] = len(group_fields)
                group_fields.append(out['field'])
            stats.append(out['stats'])

        def flush():
            for plan, _ in pending:
                plan.run()
            for plan, idx in pending:
                st = plan.sync()
                out = plan.fetch()
                out['stats'] = st
                consume(out, idx)
            del pending[:]
        stats = []
        converted = {}
        reduced_on_device = False
        for idx in groups:
            first = int(idx[0])
            params = dict(base, **{k: self._overrides[k][first] for k in other_keys})

            def member_values(key):
                # one vectorised conversion of the whole override list (1M members: a Python loop would take seconds)
                if key not in converted:
                    converted[key] = np.asarray(self._overrides[key], dtype=np.float64)
                return converted[key][idx]

            def member_array(key):
                if key in self._overrides:
                    return np.ascontiguousarray(member_values(key).reshape(len(idx), N, 3))
                return np.ascontiguousarray(np.asarray(base[key], dtype=np.float64).reshape(N, 3))

            radius = params['radius']
            if N == 1 and 'radius' in self._overrides:
                radius = np.ascontiguousarray(member_values('radius').reshape(len(idx), 1))
            temperature = params['temperature']
            if N == 1 and 'temperature' in self._overrides:
                temperature = np.ascontiguousarray(member_values('temperature').reshape(len(idx)))

            def member_scalar(key, column=False):
                if key in self._overrides and key in fast:
                    v = member_values(key).reshape(len(idx))
                    return np.ascontiguousarray(v.reshape(len(idx), 1) if column else v)
                return params[key]
            args = (radius, member_scalar('anisotropy', True), member_array('anisotropy_axis'),
                    member_array('magnetisation_direction'), params['location'], params['magnetisation'],
                    member_scalar('damping'), temperature, renorm, interactions, implicit_solve, time_step,
                    end_time, S, seeds[lo:hi] if single else seeds[idx], params['field_shape'],
                    member_scalar('field_amplitude'), params['field_frequency'], implicit_tol)
            kwargs = dict(device=device, return_trajectories=return_trajectories,
                          return_sums=True, return_final=True, gauss=gauss, implicit_newton=implicit_newton)
            if single:
                kwargs['stream_offset'] = int(stream_offset) + lo   # contiguous slice: global index = offset + position
                if device_comm is not None and devices is None:
                    kwargs['comm'] = device_comm                   # sums all-reduced on the device, inside the call
                    reduced_on_device = True
            else:
                # a parameter group is a scattered subset: every member keeps its GLOBAL index in the Philox counter,
                # whatever the grouping or the sharding
                kwargs['member_index'] = np.asarray(idx, dtype=np.uint64) + np.uint64(int(stream_offset))
            if single or devices is not None:
                consume(core.simulate_ensemble(*args, devices=devices, **kwargs), idx)
            else:
                # several parameter groups: one plan (own CUDA stream) per group, all enqueued before the first is waited
                # for, so that the groups' kernels — each usually far too small to fill the GPU — run concurrently
                try:
                    plan = core.EnsemblePlan(*args, **kwargs)
                except MemoryError:
                    # every pending plan holds its own device buffers: run and release them, then try this group again
                    if not pending:
                        raise
                    flush()
                    plan = core.EnsemblePlan(*args, **kwargs)
                pending.append((plan, idx))
                if len(pending) >= _MAX_CONCURRENT_PLANS:
                    flush()
        flush()
        time, field, traj, final, sums = state['time'], state['field'], state['traj'], state['final'], state['sums']
        if comm is not None and not reduced_on_device:
            comm.allreduce_sums(sums)
        member_fields = None
        if member_amp is not None:
            member_fields = (np.stack(group_fields), group_of, member_amp)
        elif len(group_fields) > 1 and any(not np.array_equal(f, group_fields[0]) for f in group_fields[1:]):
            member_fields = (np.stack(group_fields), group_of, None)   # members see different fields: keep each one's own
        return EnsembleResults.from_arrays(time, field, R, trajectories=traj, sums=sums, final=final, stats=stats,
                                           member_fields=member_fields)


class DOModel:
    """Discrete-orientation (two-state master equation) model of a single uniaxial particle in a field along its
    axis — `magpy.DOModel` (magpy/model.py:211-296): same constructor arguments, `simulate(end_time, time_step,
    max_samples)` returns a `Results` whose z magnetisation is the unitless p_0 - p_1 the reference returns.
    `field_shape` may be 'constant', 'sine', 'square' or 'square_f' (`field_n_components` Fourier terms).

    `simulate_batch` integrates many particles in one device call (one GPU thread each): pass arrays of radii and
    anisotropy constants (e.g. a size distribution) — the case a GPU is for; the reference has no equivalent."""

    def __init__(self, radius, anisotropy, initial_probabilities, magnetisation, damping, temperature,
                 field_shape='constant', field_frequency=0.0, field_amplitude=0.0, field_n_components=1):
        self.radius = radius
        self.volume = 4. / 3 * np.pi * self.radius ** 3
        self.anisotropy = anisotropy
        self.initial_probabilities = np.array(initial_probabilities, dtype=np.float64)
        self.magnetisation = magnetisation
        self.damping = damping
        self.temperature = temperature
        self.field_shape = field_shape
        self.field_frequency = field_frequency
        self.field_amplitude = field_amplitude
        self.field_n_components = field_n_components

    def simulate(self, end_time, time_step, max_samples):
        results = core.simulate_dom(
            np.ascontiguousarray(self.initial_probabilities, dtype=np.float64), float(self.volume), float(self.anisotropy),
            float(self.temperature), float(self.magnetisation), float(self.damping), time_step, end_time, max_samples,
            self.field_shape, self.field_amplitude, self.field_frequency, self.field_n_components)
        return Results(**results)

    def simulate_batch(self, radius, anisotropy, end_time, time_step, max_samples, initial_probabilities=None, device=0):
        """The same model over arrays `radius` and `anisotropy` (n,): {'time', 'field' (n, S), 'mz' (n, S), 'steps'}."""
        radius = np.atleast_1d(np.asarray(radius, dtype=np.float64))
        anisotropy = np.ascontiguousarray(np.broadcast_to(np.asarray(anisotropy, dtype=np.float64), radius.shape))
        p0 = self.initial_probabilities if initial_probabilities is None else np.asarray(initial_probabilities, dtype=np.float64)
        p0 = np.ascontiguousarray(np.broadcast_to(p0, (len(radius), 2)))
        volume = 4. / 3 * np.pi * radius ** 3
        return core.simulate_dom_batch(p0, volume, anisotropy, float(self.temperature), float(self.magnetisation),
                                       float(self.damping), time_step, end_time, max_samples, self.field_shape,
                                       self.field_amplitude, self.field_frequency, self.field_n_components, device)
