# cython: language_level=3, boundscheck=False, wraparound=False
"""magpy_b200.core — Cython layer over the C ABI of libmagpy_b200 (include/magpy_b200.h).

Drop-in for the sLLG entry points of the reference's ``magpy/core.pyx``:

* ``simulate(...)`` keeps the exact Python signature, defaults and return dict of
  ``magpy/core.pyx:128-203`` (which binds ``simulation::full_dynamics``,
  ``magpy/core.pyx:72-93``); a bad ``field_shape`` still raises ``KeyError``
  (``magpy/core.pyx:158-161``).
* ``get_KB / get_mu0 / get_gamma`` as ``magpy/core.pyx:29-34``.
* ``simulate_ensemble(...)`` / ``EnsemblePlan`` are the batched entry that replaces the
  per-member joblib fan-out of ``magpy/model.py:202-207``.

The GIL is released around every library call.  There is no CPU fallback: without a
CUDA device the library returns an error and this module raises ``RuntimeError``.
"""
import numpy as np
cimport numpy as np
from libc.stdint cimport int32_t, int64_t, uint32_t, uint64_t
from libc.string cimport memset

np.import_array()

cdef extern from "magpy_b200.h" nogil:
    int MAGPY_B200_ABI_VERSION
    int MAGPY_B200_OK
    int MAGPY_B200_ERR_BAD_ARG
    int MAGPY_B200_ERR_NO_DEVICE
    int MAGPY_B200_ERR_CUDA
    int MAGPY_B200_ERR_NOMEM
    int MAGPY_B200_FIELD_SINE
    int MAGPY_B200_FIELD_SQUARE
    int MAGPY_B200_FIELD_CONSTANT
    int MAGPY_B200_GAUSS_F32
    int MAGPY_B200_GAUSS_F64
    int MAGPY_B200_GAUSS_F32_PACKED

    ctypedef struct magpy_b200_stats:
        uint64_t steps_per_member
        uint64_t particle_steps
        uint64_t newton_iterations
        uint64_t newton_max_iterations
        uint64_t newton_failures
        uint64_t kernel_launches
        double device_ms
        double integrate_ms
        uint64_t h2d_bytes
        uint64_t d2h_bytes
        uint64_t kernel_family
        uint64_t kernel_variant
        double host_setup_ms
        double host_run_ms
        double host_fetch_ms
        double host_total_ms

    ctypedef struct magpy_b200_ensemble:
        uint32_t abi_version
        int32_t device
        uint64_t n_members
        uint32_t n_particles
        const double* radius
        const double* anisotropy
        const double* location
        const double* anisotropy_axis
        uint64_t axis_stride
        const double* magnetisation_direction
        uint64_t m0_stride
        double magnetisation
        double damping
        double temperature
        int32_t renorm
        int32_t interactions
        int32_t use_implicit
        double implicit_tol
        double time_step
        double end_time
        uint64_t max_samples
        int32_t field_shape
        double field_amplitude
        double field_frequency
        const int64_t* seeds
        uint64_t stream_offset
        int32_t gauss_mode
        const double* injected_dw
        uint64_t injected_steps
        double* out_time
        double* out_field
        double* out_trajectories
        double* out_sums
        double* out_final
        uint32_t noise_coarsen_log2
        uint32_t implicit_newton
        uint64_t radius_stride
        const double* member_temperature
        const uint64_t* member_index
        const double* member_anisotropy
        const double* member_damping
        const double* member_field_amplitude
        magpy_b200_comm* comm

    ctypedef struct magpy_b200_plan:
        pass

    ctypedef struct magpy_b200_comm:
        pass

    int MAGPY_B200_ERR_COMM
    int magpy_b200_host_alloc(size_t bytes, void** ptr)
    int magpy_b200_host_free(void* ptr)
    int magpy_b200_host_cache_release()
    int magpy_b200_comm_unique_id(unsigned char* id)
    int magpy_b200_comm_create(const unsigned char* id, int rank, int world_size, int device, magpy_b200_comm** comm)
    int magpy_b200_comm_create_from_env(int device, magpy_b200_comm** comm)
    int magpy_b200_comm_rank(const magpy_b200_comm* comm, int* rank, int* world_size)
    int magpy_b200_comm_allreduce(magpy_b200_comm* comm, double* host_values, size_t n, int op)
    int magpy_b200_comm_barrier(magpy_b200_comm* comm)
    int magpy_b200_comm_destroy(magpy_b200_comm* comm)

    int magpy_b200_abi_version()
    const char* magpy_b200_last_error()
    int magpy_b200_device_count(int* count)
    int magpy_b200_release_cached_memory(int device)
    double magpy_b200_get_KB()
    double magpy_b200_get_mu0()
    double magpy_b200_get_gamma()
    int magpy_b200_simulate(const double*, const double*, const double*, const double*, const double*, size_t,
                            double, double, double, int, int, int, double, double, double, size_t, int64_t, int,
                            double, double, double*, double*, double*, magpy_b200_stats*)
    int magpy_b200_simulate_ensemble(const magpy_b200_ensemble*, magpy_b200_stats*)
    int magpy_b200_simulate_ensemble_multi(const magpy_b200_ensemble*, const int*, int, magpy_b200_stats*)
    int magpy_b200_plan_create(const magpy_b200_ensemble*, magpy_b200_plan**)
    int magpy_b200_plan_run(magpy_b200_plan*)
    int magpy_b200_plan_sync(magpy_b200_plan*, magpy_b200_stats*)
    int magpy_b200_plan_fetch(magpy_b200_plan*, double*, double*, double*, double*, double*)
    int magpy_b200_plan_sums_device_ptr(magpy_b200_plan*, void**, size_t*)
    int magpy_b200_plan_destroy(magpy_b200_plan*)
    int magpy_b200_reduce_units(const double*, const double*, size_t, double, double, double, double, double,
                                double, double, double*, double*, double*, double*)
    int magpy_b200_schedule(double, double, size_t, uint64_t*)
    int magpy_b200_philox_words(int, const uint32_t*, const uint32_t*, uint32_t*)
    int magpy_b200_gaussians(int, int64_t, uint64_t, uint32_t, uint64_t, uint64_t, int, double*)
    int magpy_b200_solve3(int, size_t, const double*, const double*, double*, int*)
    int magpy_b200_gaussian_stats(int, int64_t, uint64_t, uint64_t, uint64_t, int, uint64_t*, uint64_t*, double*)
    int magpy_b200_fp64_peak(int, double*, double*)
    int magpy_b200_fp64_mma_peak(int, double*)
    int magpy_b200_simulate_dom(int, size_t, const double*, const double*, const double*, double, double, double, double,
                                double, size_t, int, double, double, size_t, double*, double*, double*, uint64_t*)


cdef extern from "Python.h":
    void Py_INCREF(object)


cdef class _PinnedBlock:
    """Owner of one page-locked host block (magpy_b200_host_alloc); returns it to the library's cache when the numpy
    array built on it is garbage collected."""
    cdef void* p

    def __cinit__(self):
        self.p = NULL

    def __dealloc__(self):
        if self.p != NULL:
            magpy_b200_host_free(self.p)
            self.p = NULL


import os as _os
_PINNED_MIN_BYTES = 1 << 20


cdef object _output_array(tuple shape):
    """float64 output array for a device->host copy: page-locked (one DMA transfer, no first-touch page faults) when it
    is large enough to matter and pinned memory is available, else a plain numpy.empty.  MAGPY_B200_PINNED_OUTPUTS=0
    switches the pinned path off."""
    cdef size_t n = 8
    for d in shape:
        n *= <size_t> d
    if n < _PINNED_MIN_BYTES or _os.environ.get('MAGPY_B200_PINNED_OUTPUTS', '1') == '0':
        return np.empty(shape)
    cdef void* p = NULL
    if magpy_b200_host_alloc(n, &p) != 0 or p == NULL:
        return np.empty(shape)
    cdef _PinnedBlock blk = _PinnedBlock()
    blk.p = p
    cdef np.npy_intp dims[8]
    cdef int nd = len(shape)
    for i in range(nd):
        dims[i] = shape[i]
    cdef np.ndarray arr = np.PyArray_SimpleNewFromData(nd, dims, np.NPY_FLOAT64, p)
    Py_INCREF(blk)
    np.PyArray_SetBaseObject(arr, blk)
    return arr


def release_pinned_cache():
    """Return the cached page-locked output blocks to the system."""
    magpy_b200_host_cache_release()


# field::options numbering (include/field.hpp:95-97, magpy/core.pyx:38-42)
SINE = 0
SQUARE = 1
CONSTANT = 2

_FIELD_LOOKUP = {'constant': CONSTANT, 'sine': SINE, 'square': SQUARE}
_GAUSS_LOOKUP = {'f32': 0, 'f64': 1, 'f32p': 2}
_NEWTON_LOOKUP = {'reference': 0, 'exact': 1}


cdef _raise(int rc):
    msg = magpy_b200_last_error().decode('utf-8', 'replace')
    if rc == MAGPY_B200_ERR_BAD_ARG:
        raise ValueError(msg)
    if rc == MAGPY_B200_ERR_NOMEM:
        raise MemoryError(msg)
    if rc == MAGPY_B200_ERR_COMM:
        raise ConnectionError(msg)
    raise RuntimeError(msg)


_KERNEL_NAMES = {1: 'heun_single', 2: 'imid_single', 3: 'heun_small', 4: 'imid_small', 5: 'heun_cluster',
                 6: 'imid_cluster', 7: 'heun_cluster_mma', 8: 'imid_split', 9: 'imid_cluster_mma', 10: 'heun_cluster_big',
                 11: 'imid_cluster_big', 12: 'imid_warps'}


cdef dict _stats_dict(magpy_b200_stats* st):
    return {
        'steps_per_member': st.steps_per_member,
        'particle_steps': st.particle_steps,
        'newton_iterations': st.newton_iterations,
        'newton_max_iterations': st.newton_max_iterations,
        'newton_failures': st.newton_failures,
        'kernel_launches': st.kernel_launches,
        'device_ms': st.device_ms,
        'integrate_ms': st.integrate_ms,
        'h2d_bytes': st.h2d_bytes,
        'd2h_bytes': st.d2h_bytes,
        'kernel': _KERNEL_NAMES.get(st.kernel_family, 'unknown'),
        'kernel_variant': st.kernel_variant,
        'host_setup_ms': st.host_setup_ms, 'host_run_ms': st.host_run_ms, 'host_fetch_ms': st.host_fetch_ms,
        'host_total_ms': st.host_total_ms,
    }


cdef class Comm:
    """One rank of a one-process-per-GPU job: the library's NCCL communicator (include/magpy_b200.h, multi-GPU section).

    `Comm.from_env(device=-1)` follows the launcher's environment (RANK, WORLD_SIZE, LOCAL_RANK, MASTER_ADDR,
    MASTER_PORT — torchrun's contract); `Comm(id, rank, world_size, device)` takes an id from `Comm.unique_id()` that the
    caller distributed itself.  Pass it as `comm=` to `simulate_ensemble` / `EnsemblePlan` /
    `EnsembleModel.simulate(shard=...)`: the ensemble sums are then all-reduced on the device (ONE ncclAllReduce per
    pass)."""
    cdef magpy_b200_comm* c
    cdef readonly int rank, world_size

    def __cinit__(self):
        self.c = NULL

    def __init__(self, bytes id=None, int rank=0, int world_size=1, int device=0):
        cdef const unsigned char* p = NULL
        if id is not None:
            if len(id) != 128:
                raise ValueError('id must be the 128 bytes of Comm.unique_id()')
            p = <const unsigned char*> id
        cdef int rc
        with nogil:
            rc = magpy_b200_comm_create(p, rank, world_size, device, &self.c)
        if rc != 0:
            self.c = NULL
            _raise(rc)
        self.rank, self.world_size = rank, world_size

    @staticmethod
    def unique_id():
        cdef unsigned char buf[128]
        cdef int rc = magpy_b200_comm_unique_id(buf)
        if rc != 0:
            _raise(rc)
        return bytes(buf[:128])

    @staticmethod
    def from_env(int device=-1):
        cdef Comm self = Comm.__new__(Comm)
        cdef int rc
        with nogil:
            rc = magpy_b200_comm_create_from_env(device, &self.c)
        if rc != 0:
            self.c = NULL
            _raise(rc)
        magpy_b200_comm_rank(self.c, &self.rank, &self.world_size)
        return self

    def __dealloc__(self):
        if self.c != NULL:
            magpy_b200_comm_destroy(self.c)
            self.c = NULL

    def allreduce(self, values, str op='sum'):
        """In-place all-reduce of a small host float64 array over all ranks ('sum' or 'max')."""
        cdef np.ndarray[double, ndim=1, mode='c'] v = np.ascontiguousarray(values, dtype=np.float64).reshape(-1)
        cdef int code = {'sum': 0, 'max': 1}[op]
        cdef size_t n = v.shape[0]
        cdef int rc
        cdef double* p = &v[0] if n else NULL
        with nogil:
            rc = magpy_b200_comm_allreduce(self.c, p, n, code)
        if rc != 0:
            _raise(rc)
        out = np.asarray(values)
        if isinstance(values, np.ndarray):
            values[...] = v.reshape(values.shape)
            return values
        return v.reshape(out.shape)

    def allreduce_sums(self, sums):
        """Sum the [S][4] ensemble sums over all ranks (host array)."""
        return self.allreduce(sums, 'sum')

    def barrier(self):
        cdef int rc
        with nogil:
            rc = magpy_b200_comm_barrier(self.c)
        if rc != 0:
            _raise(rc)


cpdef get_KB():
    return magpy_b200_get_KB()

cpdef get_mu0():
    return magpy_b200_get_mu0()

cpdef get_gamma():
    return magpy_b200_get_gamma()

cpdef int abi_version():
    return magpy_b200_abi_version()

cpdef int device_count():
    cdef int n = 0
    magpy_b200_device_count(&n)
    return n


def release_cached_memory(int device=0):
    """Return the device buffers cached from earlier calls to the CUDA driver."""
    cdef int rc = magpy_b200_release_cached_memory(device)
    if rc != 0:
        _raise(rc)


cpdef simulate(
    np.ndarray[double, ndim=1, mode='c'] radius,
    np.ndarray[double, ndim=1, mode='c'] anisotropy,
    np.ndarray[double, ndim=2, mode='c'] anisotropy_axis,
    np.ndarray[double, ndim=2, mode='c'] magnetisation_direction,
    np.ndarray[double, ndim=2, mode='c'] location,
    double magnetisation,
    double damping,
    double temperature,
    bint renorm,
    bint interactions,
    bint use_implicit,
    double time_step,
    double end_time,
    int max_samples,
    long seed,
    str field_shape='constant',
    double field_amplitude=0.0,
    double field_frequency=0.0,
    double implicit_tol=1e-9 ):
    """Same signature and return dict as the reference's ``core.simulate``
    (magpy/core.pyx:128-203); the noise is the in-kernel Philox stream keyed by `seed`."""
    cdef size_t n_particles = radius.shape[0]
    if (anisotropy.shape[0] != n_particles or anisotropy_axis.shape[0] != n_particles
            or magnetisation_direction.shape[0] != n_particles or location.shape[0] != n_particles
            or anisotropy_axis.shape[1] != 3 or magnetisation_direction.shape[1] != 3 or location.shape[1] != 3):
        raise ValueError('per-particle arrays must have matching lengths and 3 components')
    if max_samples < 2:
        raise ValueError('max_samples must be >= 2')
    cdef int field_code = _FIELD_LOOKUP[field_shape]   # KeyError on a bad shape, as the reference
    cdef np.ndarray[double, ndim=1] t = np.empty(max_samples)
    cdef np.ndarray[double, ndim=1] fld = np.empty(max_samples)
    cdef np.ndarray[double, ndim=3] m = np.empty((n_particles, 3, max_samples))
    cdef magpy_b200_stats st
    cdef int rc
    cdef const double* p_r = &radius[0]
    cdef const double* p_k = &anisotropy[0]
    cdef const double* p_ax = &anisotropy_axis[0, 0]
    cdef const double* p_m0 = &magnetisation_direction[0, 0]
    cdef const double* p_loc = &location[0, 0]
    cdef double* p_t = &t[0]
    cdef double* p_f = &fld[0]
    cdef double* p_m = &m[0, 0, 0]
    cdef int c_renorm = renorm, c_inter = interactions, c_impl = use_implicit
    cdef size_t c_S = max_samples
    cdef int64_t c_seed = seed
    with nogil:
        rc = magpy_b200_simulate(p_r, p_k, p_ax, p_m0, p_loc, n_particles, magnetisation, damping, temperature,
                                 c_renorm, c_inter, c_impl, implicit_tol, time_step, end_time, c_S, c_seed,
                                 field_code, field_amplitude, field_frequency, p_t, p_f, p_m, &st)
    if rc != 0:
        _raise(rc)
    return {
        'N': n_particles,
        'time': t,
        'field': fld,
        'x': {i: m[i, 0] for i in range(n_particles)},
        'y': {i: m[i, 1] for i in range(n_particles)},
        'z': {i: m[i, 2] for i in range(n_particles)},
    }


cdef class _EnsembleArgs:
    """Owns the numpy buffers a magpy_b200_ensemble points into."""
    cdef magpy_b200_ensemble a
    cdef object keep
    cdef public object time, field, trajectories, sums, final
    cdef public size_t R, N, S

    def __cinit__(self):
        memset(&self.a, 0, sizeof(magpy_b200_ensemble))
        self.keep = []


cdef _EnsembleArgs _build_args(radius, anisotropy, anisotropy_axis, magnetisation_direction, location,
                               double magnetisation, damping, temperature, bint renorm,
                               bint interactions, bint use_implicit, double time_step, double end_time,
                               max_samples, seeds, str field_shape, field_amplitude,
                               double field_frequency, double implicit_tol, int device, stream_offset,
                               bint return_trajectories, bint return_sums, bint return_final, str gauss,
                               injected_dw, int noise_coarsen_log2=0, str implicit_newton='reference',
                               member_index=None, Comm comm=None):
    cdef _EnsembleArgs e = _EnsembleArgs()
    # radius: (N,) shared by all members, or (R, N) per member (single-particle ensembles: a size distribution)
    rad = np.ascontiguousarray(radius, dtype=np.float64)
    cdef bint member_radii = rad.ndim == 2
    cdef np.ndarray[double, ndim=1, mode='c'] c_radius = rad.reshape(-1)
    cdef size_t N = rad.shape[1] if member_radii else c_radius.shape[0]
    # anisotropy: (N,) shared, or (R, 1) per member (single-particle ensembles: an anisotropy distribution)
    anis = np.ascontiguousarray(anisotropy, dtype=np.float64)
    cdef bint member_anis = anis.ndim == 2
    cdef np.ndarray[double, ndim=1, mode='c'] c_anis = anis.reshape(-1)
    cdef np.ndarray[double, ndim=2, mode='c'] c_loc = np.ascontiguousarray(location, dtype=np.float64).reshape(-1, 3)
    cdef np.ndarray[np.int64_t, ndim=1, mode='c'] c_seeds = np.ascontiguousarray(seeds, dtype=np.int64).reshape(-1)
    cdef size_t R = c_seeds.shape[0]
    if R == 0:
        raise ValueError('seeds must hold one seed per ensemble member')
    if N == 0 or (not member_anis and c_anis.shape[0] != N) or c_loc.shape[0] != N:
        raise ValueError('radius, anisotropy and location must describe the same number of particles')
    if member_radii and (rad.shape[0] != R or N != 1):
        raise ValueError('per-member radii need shape (R, 1): supported for single-particle ensembles only')
    if member_anis and (anis.shape[0] != R or anis.shape[1] != 1 or N != 1):
        raise ValueError('per-member anisotropy needs shape (R, 1): supported for single-particle ensembles only')
    if int(max_samples) < 2:
        raise ValueError('max_samples must be >= 2')
    cdef size_t S = int(max_samples)
    ax = np.ascontiguousarray(anisotropy_axis, dtype=np.float64)
    m0 = np.ascontiguousarray(magnetisation_direction, dtype=np.float64)
    for name, arr in (('anisotropy_axis', ax), ('magnetisation_direction', m0)):
        if arr.shape not in ((N, 3), (R, N, 3)):
            raise ValueError('%s must have shape (N,3) or (R,N,3), got %r' % (name, arr.shape))
    cdef np.ndarray[double, ndim=1, mode='c'] c_ax = ax.reshape(-1)
    cdef np.ndarray[double, ndim=1, mode='c'] c_m0 = m0.reshape(-1)
    cdef np.ndarray[double, ndim=1, mode='c'] c_dw
    e.keep = [c_radius, c_anis, c_loc, c_seeds, c_ax, c_m0]
    e.R, e.N, e.S = R, N, S

    e.a.abi_version = MAGPY_B200_ABI_VERSION
    e.a.device = device
    e.a.n_members = R
    e.a.n_particles = N
    e.a.radius = &c_radius[0]
    e.a.radius_stride = N if member_radii else 0
    e.a.anisotropy = &c_anis[0]
    if member_anis:
        e.a.member_anisotropy = &c_anis[0]
    e.a.location = &c_loc[0, 0]
    e.a.anisotropy_axis = &c_ax[0]
    e.a.axis_stride = 3 * N if ax.ndim == 3 else 0
    e.a.magnetisation_direction = &c_m0[0]
    e.a.m0_stride = 3 * N if m0.ndim == 3 else 0
    e.a.magnetisation = magnetisation
    # damping / field_amplitude: scalars, or one value per member (single-particle ensembles)
    cdef np.ndarray[double, ndim=1, mode='c'] c_damp, c_famp
    if np.ndim(damping) == 0:
        e.a.damping = float(damping)
    else:
        c_damp = np.ascontiguousarray(damping, dtype=np.float64).reshape(-1)
        if c_damp.shape[0] != R or N != 1:
            raise ValueError('per-member damping needs one value per member: supported for single-particle ensembles only')
        e.keep.append(c_damp)
        e.a.damping = c_damp[0]
        e.a.member_damping = &c_damp[0]
    if np.ndim(field_amplitude) == 0:
        e.a.field_amplitude = float(field_amplitude)
    else:
        c_famp = np.ascontiguousarray(field_amplitude, dtype=np.float64).reshape(-1)
        if c_famp.shape[0] != R or N != 1:
            raise ValueError('per-member field amplitudes need one value per member: supported for single-particle ensembles only')
        e.keep.append(c_famp)
        e.a.field_amplitude = c_famp[0]
        e.a.member_field_amplitude = &c_famp[0]
    cdef np.ndarray[np.uint64_t, ndim=1, mode='c'] c_midx
    if member_index is not None:
        c_midx = np.ascontiguousarray(member_index, dtype=np.uint64).reshape(-1)
        if c_midx.shape[0] != R:
            raise ValueError('member_index must hold one global index per member')
        e.keep.append(c_midx)
        e.a.member_index = <const uint64_t*> &c_midx[0]
    if comm is not None:
        e.keep.append(comm)
        e.a.comm = comm.c
    # temperature: a scalar, or one value per member (single-particle ensembles: a temperature sweep in one launch)
    cdef np.ndarray[double, ndim=1, mode='c'] c_temp
    if np.ndim(temperature) == 0:
        e.a.temperature = float(temperature)
    else:
        c_temp = np.ascontiguousarray(temperature, dtype=np.float64).reshape(-1)
        if c_temp.shape[0] != R or N != 1:
            raise ValueError('per-member temperatures need one value per member: supported for single-particle ensembles only')
        e.keep.append(c_temp)
        e.a.temperature = c_temp[0]
        e.a.member_temperature = &c_temp[0]
    e.a.renorm = renorm
    e.a.interactions = interactions
    e.a.use_implicit = use_implicit
    e.a.implicit_tol = implicit_tol
    e.a.time_step = time_step
    e.a.end_time = end_time
    e.a.max_samples = S
    e.a.field_shape = _FIELD_LOOKUP[field_shape]
    e.a.field_frequency = field_frequency
    e.a.seeds = <const int64_t*> &c_seeds[0]
    e.a.stream_offset = int(stream_offset)
    e.a.gauss_mode = _GAUSS_LOOKUP[gauss]
    if noise_coarsen_log2 < 0:
        raise ValueError('noise_coarsen_log2 must be >= 0')
    e.a.noise_coarsen_log2 = noise_coarsen_log2
    e.a.implicit_newton = _NEWTON_LOOKUP[implicit_newton]   # KeyError on anything but 'reference' / 'exact'
    if injected_dw is not None:
        dw = np.ascontiguousarray(injected_dw, dtype=np.float64)
        if dw.ndim != 3 or dw.shape[0] != R or dw.shape[2] != 3 * N:
            raise ValueError('injected_dw must have shape (R, steps, 3N)')
        c_dw = dw.reshape(-1)
        e.keep.append(c_dw)
        e.a.injected_dw = &c_dw[0]
        e.a.injected_steps = dw.shape[1]

    cdef np.ndarray[double, ndim=1] o_time = np.empty(S)
    cdef np.ndarray[double, ndim=1] o_field = np.empty(S)
    cdef np.ndarray[double, ndim=4] o_traj
    cdef np.ndarray[double, ndim=2] o_sums
    cdef np.ndarray[double, ndim=3] o_final
    e.time, e.field = o_time, o_field
    e.a.out_time = &o_time[0]
    e.a.out_field = &o_field[0]
    if return_trajectories:
        o_traj = _output_array((R, N, 3, S))
        e.trajectories = o_traj
        e.a.out_trajectories = &o_traj[0, 0, 0, 0]
    if return_sums:
        o_sums = np.empty((S, 4))
        e.sums = o_sums
        e.a.out_sums = &o_sums[0, 0]
    if return_final:
        o_final = _output_array((R, N, 3))
        e.final = o_final
        e.a.out_final = &o_final[0, 0, 0]
    return e


def simulate_ensemble(radius, anisotropy, anisotropy_axis, magnetisation_direction, location,
                      double magnetisation, damping, temperature, bint renorm, bint interactions,
                      bint use_implicit, double time_step, double end_time, max_samples, seeds,
                      str field_shape='constant', field_amplitude=0.0, double field_frequency=0.0,
                      double implicit_tol=1e-9, int device=0, stream_offset=0, bint return_trajectories=True,
                      bint return_sums=True, bint return_final=True, str gauss='f32p', injected_dw=None, devices=None,
                      int noise_coarsen_log2=0, str implicit_newton='reference', member_index=None, Comm comm=None):
    """Integrate R = len(seeds) independent members of one cluster in a single call.

    `devices` (list of CUDA ordinals, or 'all') shards the members over several GPUs of this box from this one
    process (magpy_b200_simulate_ensemble_multi); `device` is then ignored.

    `noise_coarsen_log2` = L > 0 (single-particle ensembles, gauss='f32p'): the increment of step s is the
    normalised sum of the 2^L increments the L = 0 stream gives to steps s 2^L ... (s+1) 2^L - 1, so runs with
    time_step * 2^L see the same Brownian paths (convergence studies, test/convergence/task5.cpp:150-158).

    `anisotropy_axis` and `magnetisation_direction` are (N,3) (shared) or (R,N,3).  Single-particle ensembles also
    take per-member `radius` (R,1), `anisotropy` (R,1), `temperature` (R,), `damping` (R,) and `field_amplitude` (R,).
    `member_index` (R,) gives the members' global indices (word 1 of their Philox counters) when they are not the
    contiguous range starting at `stream_offset`.  `comm` (a `Comm`) all-reduces the sums over all ranks on the device.
    Returns a dict with 'time' [S], 'field' [S], 'trajectories' [R,N,3,S] | None,
    'sums' [S,4] | None (sum over members of cluster Mx,My,Mz and Mz^2), 'final' [R,N,3] | None
    and 'stats'.  `injected_dw` (R, steps, 3N) replaces the Philox stream by caller-supplied
    unit-variance increments (the reference's RngArray hook, lib/rng.cpp:67-94)."""
    cdef _EnsembleArgs e = _build_args(radius, anisotropy, anisotropy_axis, magnetisation_direction, location,
                                       magnetisation, damping, temperature, renorm, interactions, use_implicit,
                                       time_step, end_time, max_samples, seeds, field_shape, field_amplitude,
                                       field_frequency, implicit_tol, device, stream_offset, return_trajectories,
                                       return_sums, return_final, gauss, injected_dw, noise_coarsen_log2, implicit_newton,
                                       member_index, comm)
    cdef magpy_b200_stats st
    cdef int rc
    cdef np.ndarray[int, ndim=1, mode='c'] c_dev
    cdef int n_dev
    if devices is None:
        with nogil:
            rc = magpy_b200_simulate_ensemble(&e.a, &st)
    else:
        if isinstance(devices, str):
            if devices != 'all':
                raise ValueError("devices must be a list of CUDA device ordinals or 'all'")
            devices = list(range(device_count()))
            if not devices:
                raise RuntimeError('no CUDA device available; magpy_b200 has no CPU fallback')
        c_dev = np.ascontiguousarray(devices, dtype=np.intc).reshape(-1)
        n_dev = c_dev.shape[0]
        if n_dev == 0:
            raise ValueError('devices must list at least one CUDA device')
        with nogil:
            rc = magpy_b200_simulate_ensemble_multi(&e.a, <const int*> &c_dev[0], n_dev, &st)
    if rc != 0:
        _raise(rc)
    return {'N': e.N, 'R': e.R, 'time': e.time, 'field': e.field, 'trajectories': e.trajectories,
            'sums': e.sums, 'final': e.final, 'stats': _stats_dict(&st)}


cdef class EnsemblePlan:
    """Device-resident ensemble: inputs are uploaded once, `run()` re-integrates from the
    initial state on the plan's CUDA stream, `sync()` waits and returns the stats (CUDA-event
    times), `fetch()` downloads the outputs requested at construction."""
    cdef magpy_b200_plan* plan
    cdef _EnsembleArgs e

    def __cinit__(self):
        self.plan = NULL

    def __init__(self, radius, anisotropy, anisotropy_axis, magnetisation_direction, location,
                 double magnetisation, damping, temperature, bint renorm, bint interactions,
                 bint use_implicit, double time_step, double end_time, max_samples, seeds,
                 str field_shape='constant', field_amplitude=0.0, double field_frequency=0.0,
                 double implicit_tol=1e-9, int device=0, stream_offset=0, bint return_trajectories=False,
                 bint return_sums=True, bint return_final=True, str gauss='f32p', injected_dw=None,
                 str implicit_newton='reference', member_index=None, Comm comm=None):
        self.e = _build_args(radius, anisotropy, anisotropy_axis, magnetisation_direction, location,
                             magnetisation, damping, temperature, renorm, interactions, use_implicit,
                             time_step, end_time, max_samples, seeds, field_shape, field_amplitude,
                             field_frequency, implicit_tol, device, stream_offset, return_trajectories,
                             return_sums, return_final, gauss, injected_dw, 0, implicit_newton, member_index, comm)
        cdef int rc
        with nogil:
            rc = magpy_b200_plan_create(&self.e.a, &self.plan)
        if rc != 0:
            self.plan = NULL
            _raise(rc)

    def __dealloc__(self):
        if self.plan != NULL:
            magpy_b200_plan_destroy(self.plan)
            self.plan = NULL

    def run(self):
        cdef int rc
        with nogil:
            rc = magpy_b200_plan_run(self.plan)
        if rc != 0:
            _raise(rc)

    def sync(self):
        cdef magpy_b200_stats st
        cdef int rc
        with nogil:
            rc = magpy_b200_plan_sync(self.plan, &st)
        if rc != 0:
            _raise(rc)
        return _stats_dict(&st)

    def fetch(self):
        cdef int rc
        with nogil:
            rc = magpy_b200_plan_fetch(self.plan, self.e.a.out_time, self.e.a.out_field,
                                       self.e.a.out_trajectories, self.e.a.out_sums, self.e.a.out_final)
        if rc != 0:
            _raise(rc)
        return {'N': self.e.N, 'R': self.e.R, 'time': self.e.time, 'field': self.e.field,
                'trajectories': self.e.trajectories, 'sums': self.e.sums, 'final': self.e.final}

    def sums_device_ptr(self):
        """(device address, number of doubles) of the [S][4] ensemble-sum buffer (reduced units)."""
        cdef void* p = NULL
        cdef size_t n = 0
        cdef int rc = magpy_b200_plan_sums_device_ptr(self.plan, &p, &n)
        if rc != 0:
            _raise(rc)
        return <size_t> p, n


def reduce_units(radius, anisotropy, double magnetisation, double damping, double temperature,
                 double time_step, double end_time, double field_amplitude=0.0, double field_frequency=0.0):
    """SI -> reduced units exactly as lib/simulation.cpp:498-549."""
    cdef np.ndarray[double, ndim=1, mode='c'] r = np.ascontiguousarray(radius, dtype=np.float64).reshape(-1)
    cdef np.ndarray[double, ndim=1, mode='c'] k = np.ascontiguousarray(anisotropy, dtype=np.float64).reshape(-1)
    cdef size_t N = r.shape[0]
    cdef np.ndarray[double, ndim=1] k_red = np.empty(N), v_red = np.empty(N), sigma = np.empty(N), sc = np.empty(9)
    cdef int rc = magpy_b200_reduce_units(&r[0], &k[0], N, magnetisation, damping, temperature, time_step,
                                          end_time, field_amplitude, field_frequency, &k_red[0], &v_red[0],
                                          &sigma[0], &sc[0])
    if rc != 0:
        _raise(rc)
    names = ('V_av', 'K_av', 'H_k', 'time_factor', 'dt_red', 'T_red', 'h0', 'f_red', 'dipolar_prefactor')
    out = dict(zip(names, sc.tolist()))
    out.update(k_red=k_red, v_red=v_red, sigma=sigma)
    return out


def schedule(double dt_red, double t_end_red, max_samples):
    """Cumulative step count per sample of the zero-order-hold schedule (lib/simulation.cpp:342-405)."""
    cdef size_t S = int(max_samples)
    cdef np.ndarray[np.uint64_t, ndim=1] cum = np.zeros(S, dtype=np.uint64)
    cdef int rc = magpy_b200_schedule(dt_red, t_end_red, S, <uint64_t*> &cum[0])
    if rc != 0:
        _raise(rc)
    return cum


def philox_words(ctr, key, int device=0):
    cdef uint32_t c[4]
    cdef uint32_t k[2]
    cdef uint32_t o[4]
    for i in range(4):
        c[i] = ctr[i]
    k[0], k[1] = key[0], key[1]
    cdef int rc = magpy_b200_philox_words(device, c, k, o)
    if rc != 0:
        _raise(rc)
    return [o[0], o[1], o[2], o[3]]


def gaussians(seed, member, particle, first_step, n_steps, str gauss='f32p', int device=0):
    """The (n_steps, 3) unit-variance draws the kernels use for (seed, member, particle)."""
    cdef np.ndarray[double, ndim=2] out = np.empty((int(n_steps), 3))
    cdef int rc = magpy_b200_gaussians(device, int(seed), int(member), int(particle), int(first_step),
                                       int(n_steps), _GAUSS_LOOKUP[gauss], &out[0, 0])
    if rc != 0:
        _raise(rc)
    return out


def solve3(A, b, int device=0):
    """The implicit kernels' in-register 3x3 solve (adjugate) on the device: A (n, 3, 3), b (n, 3) -> x (n, 3), ok (n,)."""
    cdef np.ndarray[double, ndim=1, mode='c'] cA = np.ascontiguousarray(A, dtype=np.float64).reshape(-1)
    cdef np.ndarray[double, ndim=1, mode='c'] cb = np.ascontiguousarray(b, dtype=np.float64).reshape(-1)
    cdef size_t n = cb.shape[0] // 3
    if cA.shape[0] != 9 * n or n == 0:
        raise ValueError('A must be (n, 3, 3) and b (n, 3)')
    cdef np.ndarray[double, ndim=2] x = np.empty((n, 3))
    cdef np.ndarray[int, ndim=1] ok = np.empty(n, dtype=np.intc)
    cdef int rc = magpy_b200_solve3(device, n, &cA[0], &cb[0], &x[0, 0], <int*> &ok[0])
    if rc != 0:
        _raise(rc)
    return x, ok.astype(bool)


def gaussian_stats(seed, n_members, n_steps, str gauss='f32p', first_member=0, int device=0):
    """Histograms and moments of 3 * n_members * n_steps draws of the in-kernel Gaussian stream, accumulated on the device:
    {'n', 'hist' (4096 bins of width 1/256 over [-8, 8)), 'edges', 'angle_hist' (1024 bins over [-pi, pi)), 'n_angles',
    'sum', 'sum2', 'sum3', 'sum4', 'max_abs'}."""
    cdef np.ndarray[np.uint64_t, ndim=1] hist = np.zeros(4096, dtype=np.uint64)
    cdef np.ndarray[np.uint64_t, ndim=1] ang = np.zeros(1024, dtype=np.uint64)
    cdef np.ndarray[double, ndim=1] mom = np.zeros(5)
    cdef int64_t c_seed = int(seed)
    cdef uint64_t c_first = int(first_member), c_n = int(n_members), c_steps = int(n_steps)
    cdef int code = _GAUSS_LOOKUP[gauss]
    cdef int rc
    with nogil:
        rc = magpy_b200_gaussian_stats(device, c_seed, c_first, c_n, c_steps, code, <uint64_t*> &hist[0],
                                       <uint64_t*> &ang[0], &mom[0])
    if rc != 0:
        _raise(rc)
    return {'n': 3 * int(n_members) * int(n_steps), 'hist': hist, 'edges': np.arange(4097) / 256.0 - 8.0,
            'angle_hist': ang, 'n_angles': int(ang.sum()), 'sum': mom[0], 'sum2': mom[1], 'sum3': mom[2], 'sum4': mom[3],
            'max_abs': mom[4]}


def fp64_peak(int device=0):
    """Measured sustained FP64 FMA rate of the device in TFLOP/s and the max SM clock in MHz."""
    cdef double tf = 0.0, mhz = 0.0
    cdef int rc
    with nogil:
        rc = magpy_b200_fp64_peak(device, &tf, &mhz)
    if rc != 0:
        _raise(rc)
    return tf, mhz


def fp64_mma_peak(int device=0):
    """Measured sustained rate of the FP64 matrix instruction (DMMA.8x8x4) in TFLOP/s."""
    cdef double tf = 0.0
    cdef int rc
    with nogil:
        rc = magpy_b200_fp64_mma_peak(device, &tf)
    if rc != 0:
        _raise(rc)
    return tf


_DOM_FIELD_LOOKUP = {'sine': 0, 'square': 1, 'constant': 2, 'square_f': 3}


def simulate_dom_batch(np.ndarray[double, ndim=2, mode='c'] initial_probabilities,
                       np.ndarray[double, ndim=1, mode='c'] volume,
                       np.ndarray[double, ndim=1, mode='c'] anisotropy,
                       double temperature, double magnetisation, double alpha, double time_step, double end_time,
                       size_t max_samples, str field_shape='constant', double field_amplitude=0.0,
                       double field_frequency=0.0, size_t field_n_components=1, int device=0):
    """Discrete-orientation model of a BATCH of particles (one GPU thread each): `volume`, `anisotropy` (n,) and
    `initial_probabilities` (n, 2) per item, everything else shared.  Returns {'time' (S,), 'field' (n, S) in A/m,
    'mz' (n, S) = p_0 - p_1, 'steps' (n,) accepted RK45 steps}.  n = 1 is the reference's `simulate_dom`."""
    cdef size_t n = volume.shape[0]
    if anisotropy.shape[0] != n or initial_probabilities.shape[0] != n or initial_probabilities.shape[1] != 2:
        raise ValueError('volume (n,), anisotropy (n,) and initial_probabilities (n, 2) must match')
    if max_samples < 2:
        raise ValueError('max_samples must be >= 2')
    cdef int code = _DOM_FIELD_LOOKUP[field_shape]   # KeyError on a bad shape
    cdef np.ndarray[double, ndim=1] t = np.empty(max_samples)
    cdef np.ndarray[double, ndim=2] fld = np.empty((n, max_samples))
    cdef np.ndarray[double, ndim=2] mz = np.empty((n, max_samples))
    cdef np.ndarray[uint64_t, ndim=1] steps = np.empty(n, dtype=np.uint64)
    cdef const double* p_v = &volume[0]
    cdef const double* p_k = &anisotropy[0]
    cdef const double* p_p = &initial_probabilities[0, 0]
    cdef double* p_t = &t[0]
    cdef double* p_f = &fld[0, 0]
    cdef double* p_m = &mz[0, 0]
    cdef uint64_t* p_s = &steps[0]
    cdef int rc
    with nogil:
        rc = magpy_b200_simulate_dom(device, n, p_v, p_k, p_p, temperature, magnetisation, alpha, time_step, end_time,
                                     max_samples, code, field_amplitude, field_frequency, field_n_components, p_t, p_f,
                                     p_m, p_s)
    if rc != 0:
        _raise(rc)
    return {'time': t, 'field': fld, 'mz': mz, 'steps': steps}


cpdef simulate_dom(np.ndarray[double, ndim=1, mode='c'] initial_probabilities, double volume, double anisotropy,
                   double temperature, double magnetisation, double alpha, double time_step, double end_time,
                   size_t max_samples, str field_shape, double field_amplitude, double field_frequency,
                   size_t field_n_components):
    """Same signature and return dict as the reference's ``core.simulate_dom`` (magpy/core.pyx:205-278): 'z' is the
    unitless p_0 - p_1, 'x' and 'y' are zero, 'field' is in A/m."""
    out = simulate_dom_batch(np.ascontiguousarray(initial_probabilities[:2]).reshape(1, 2), np.array([volume]),
                             np.array([anisotropy]), temperature, magnetisation, alpha, time_step, end_time, max_samples,
                             field_shape, field_amplitude, field_frequency, field_n_components)
    zeros = np.zeros(max_samples)
    return {'N': 1, 'time': out['time'], 'field': out['field'][0], 'x': {0: zeros}, 'y': {0: zeros.copy()},
            'z': {0: out['mz'][0]}}
