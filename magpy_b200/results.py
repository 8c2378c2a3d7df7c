"""Results containers with the interface of the reference's ``magpy/results.py``.

``Results`` is the per-cluster container (magpy/results.py:6-91).  ``EnsembleResults``
keeps the reference's methods (magpy/results.py:94-217) but can be backed by the arrays the
batched GPU call returns — one ``[R, N, 3, S]`` trajectory block, the ``[S, 4]`` ensemble
sums reduced on the device and the ``[R, N, 3]`` final states — so a million-member
ensemble does not materialise a million Python objects.  ``results`` is then a lazy
sequence of ``Results`` views.
"""
import numpy as np

from .core import get_mu0

_DIR = {'x': 0, 'y': 1, 'z': 2}


class Results:
    """Results of a simulation of a single particle cluster (magpy/results.py:6-91).

    Args:
        time (np.ndarray): length `M`, seconds.
        field (np.ndarray): length `M`, applied field along z in A/m.
        x, y, z (dict): ``{particle id: np.ndarray of length M}`` magnetisation components.
        N (int): number of particles in the cluster.
    """
    def __init__(self, time, field, x, y, z, N):
        self.time = time
        self.field = field
        self.x = x
        self.y = y
        self.z = z
        self.N = N

    def plot(self):
        """One panel per particle with its three magnetisation components against time
        (same figure as magpy/results.py:37-59); needs matplotlib."""
        from matplotlib import pyplot
        figure, panels = pyplot.subplots(nrows=self.N, squeeze=False)
        for particle, panel in enumerate(panels[:, 0]):
            for name in 'xyz':
                panel.plot(self.time, getattr(self, name)[particle], label=name)
            panel.set_title('Particle {}'.format(particle))
            panel.set_xlabel('Reduced time [dimless]')
            panel.legend()
        figure.tight_layout()
        return figure

    def magnetisation(self, direction='z'):
        """Total (summed over particles) magnetisation along `direction` (magpy/results.py:61-75)."""
        return np.sum([vals for vals in getattr(self, direction).values()], axis=0)

    def final_state(self):
        """Last sample of every particle (magpy/results.py:77-91): {'x': {id: value}, 'y': ..., 'z': ...}."""
        return {name: {pid: series[-1] for pid, series in getattr(self, name).items()} for name in 'xyz'}


def save_results(fname, results, particle=0):
    """Write one particle of a :class:`Results` to disk in the reference's raw format
    (lib/simulation.cpp:38-63, include/io.hpp:12-27): native-endian float64 arrays, no header, in the five
    files `fname.mx`, `fname.my`, `fname.mz`, `fname.field`, `fname.time`."""
    for suffix, arr in (('mx', results.x[particle]), ('my', results.y[particle]), ('mz', results.z[particle]),
                        ('field', results.field), ('time', results.time)):
        np.ascontiguousarray(arr, dtype=np.float64).tofile('%s.%s' % (fname, suffix))


def load_results(fname):
    """Read the five raw float64 files written by :func:`save_results` (or by the reference's
    `simulation::save_results`) back into a single-particle :class:`Results`."""
    arr = {suffix: np.fromfile('%s.%s' % (fname, suffix), dtype=np.float64)
           for suffix in ('mx', 'my', 'mz', 'field', 'time')}
    return Results(arr['time'], arr['field'], {0: arr['mx']}, {0: arr['my']}, {0: arr['mz']}, 1)


_trapezoid = getattr(np, 'trapezoid', None) or np.trapz   # NumPy >= 2.0 / 1.x (the reference: scipy.integrate.trapz)


class _LazyResults:
    """Sequence of per-member ``Results`` views over one [R, N, 3, S] array.  `member_fields` = (fields [G, S],
    group index of every member [R], per-member amplitude [R] or None) when the members do not all see the same applied field (an ensemble whose
    members differ in field amplitude / frequency / shape): member i then reports its own field, as the reference's
    per-member ``Results`` do."""
    def __init__(self, time, field, traj, member_fields=None):
        self._time, self._field, self._traj = time, field, traj
        self._member_fields = member_fields

    def __len__(self):
        return self._traj.shape[0]

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        block = self._traj[i]
        N = block.shape[0]
        field = self._field
        if self._member_fields is not None:
            fields, group_of, scale = self._member_fields
            field = fields[group_of[i]]
            if scale is not None:      # per-member field amplitudes: `fields` holds the waveform for 1 A/m
                field = field * scale[i]
        return Results(self._time, field,
                       {p: block[p, 0] for p in range(N)},
                       {p: block[p, 1] for p in range(N)},
                       {p: block[p, 2] for p in range(N)}, N)

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class EnsembleResults:
    """Results of an ensemble of particle clusters (magpy/results.py:94-217).

    Construct either from a list of ``Results`` (the reference's signature) or with
    :meth:`from_arrays`.
    """
    def __init__(self, results):
        self.results = results
        self.time = results[0].time
        self.field = results[0].field
        self._traj = None
        self._sums = None
        self._final = None
        self._n_members = len(results)
        self.stats = None

    @classmethod
    def from_arrays(cls, time, field, n_members, trajectories=None, sums=None, final=None, stats=None, member_fields=None):
        self = cls.__new__(cls)
        self.time = time
        self.field = field
        self._traj = trajectories
        self._sums = sums
        self._final = final
        self._n_members = int(n_members)
        self.stats = stats
        self.results = _LazyResults(time, field, trajectories, member_fields) if trajectories is not None else None
        return self

    def __len__(self):
        return self._n_members

    def _need_traj(self, what):
        if self.results is None:
            raise ValueError(what + ' needs per-member trajectories; simulate with return_trajectories=True')

    def magnetisation(self, direction='z'):
        """Total magnetisation of each member (magpy/results.py:118-132)."""
        self._need_traj('magnetisation()')
        if self._traj is not None:
            return list(self._traj[:, :, _DIR[direction], :].sum(axis=1))
        return [res.magnetisation(direction) for res in self.results]

    def ensemble_magnetisation(self, direction='z'):
        """Mean over members of the cluster magnetisation (magpy/results.py:134-151)."""
        if self._sums is not None:
            return self._sums[:, _DIR[direction]] / self._n_members
        return np.sum(self.magnetisation(direction), axis=0) / len(self.results)

    def ensemble_magnetisation_stderr(self):
        """Standard error of the ensemble-mean z magnetisation (from the device-side sum of squares)."""
        if self._sums is None:
            mz = np.asarray(self.magnetisation('z'))
            return mz.std(axis=0, ddof=1) / np.sqrt(mz.shape[0])
        n = self._n_members
        mean = self._sums[:, 2] / n
        var = np.maximum(self._sums[:, 3] / n - mean * mean, 0.0) * (n / max(n - 1, 1))
        return np.sqrt(var / n)

    def final_state(self):
        """Final state of each member (magpy/results.py:153-165)."""
        if self._final is not None:
            f = self._final
            N = f.shape[1]
            return [{'x': {p: f[i, p, 0] for p in range(N)},
                     'y': {p: f[i, p, 1] for p in range(N)},
                     'z': {p: f[i, p, 2] for p in range(N)}} for i in range(f.shape[0])]
        self._need_traj('final_state()')
        return [res.final_state() for res in self.results]

    def final_state_array(self):
        """[R, N, 3] array of final states (A/m)."""
        if self._final is not None:
            return self._final
        self._need_traj('final_state_array()')
        return np.array([[[r.x[p][-1], r.y[p][-1], r.z[p][-1]] for p in range(r.N)] for r in self.results])

    def energy_dissipated(self, start_time=None, end_time=None):
        """Hysteresis-loop area ``-mu0 * integral H dM`` over the window (magpy/results.py:167-191)."""
        before_mask = (self.time >= start_time) if start_time is not None else True
        after_mask = (self.time <= end_time) if end_time is not None else True
        mask = before_mask & after_mask
        if mask is True:
            mask = np.ones(len(self.time), dtype=bool)
        return -get_mu0() * _trapezoid(self.field[mask], self.ensemble_magnetisation()[mask])

    def final_cycle_energy_dissipated(self, field_frequency):
        """Energy dissipated during the last field period (magpy/results.py:193-217)."""
        T = 1. / field_frequency
        return self.energy_dissipated(start_time=self.time[-1] - T)

    def specific_absorption_rate(self, field_frequency, density=5180.0):
        """Specific absorption rate in W/kg of particle material: the energy dissipated per field cycle and
        unit volume (:meth:`final_cycle_energy_dissipated`, J/m^3) times the field frequency, divided by the
        mass density (default: magnetite, 5180 kg/m^3).  The reference stops at the energy per cycle
        (magpy/results.py:193-217); the conversion is the textbook SAR = |E| f / rho.  The magnitude is
        taken because the reference's `-mu0 * trapz(field, M)` is negative for a magnetisation that lags
        the field (its sign convention is kept unchanged in :meth:`energy_dissipated`).  For a cluster the
        ensemble magnetisation is the sum over its particles, so divide by the particle count for a
        per-particle figure."""
        return abs(self.final_cycle_energy_dissipated(field_frequency)) * field_frequency / density
