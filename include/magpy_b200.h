/* magpy_b200.h — C ABI of the B200-native ensemble stochastic-LLG integrator.
 *
 * This is the drop-in boundary for ONE path of owlas/magpy: everything that
 * magpy/core.pyx reaches through `simulation::full_dynamics`.  Each entry point
 * names the reference interface it replaces (paths relative to the reference
 * repository).  Plain pointers and sizes only; the caller owns every buffer;
 * nothing allocated by the library is ever returned (except opaque plan handles
 * that must be released with magpy_b200_plan_destroy).  No C++ exception
 * crosses this boundary: every function returns an int status and
 * magpy_b200_last_error() describes the most recent failure on this thread.
 *
 * There is no CPU fallback: without a CUDA device every compute entry point
 * returns MAGPY_B200_ERR_NO_DEVICE.
 */
#ifndef MAGPY_B200_H
#define MAGPY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAGPY_B200_ABI_VERSION 4

/* status codes */
#define MAGPY_B200_OK 0
#define MAGPY_B200_ERR_BAD_ARG 1   /* invalid argument (message says which)            */
#define MAGPY_B200_ERR_NO_DEVICE 2 /* no usable CUDA device / bad device index          */
#define MAGPY_B200_ERR_CUDA 3      /* CUDA runtime error                                */
#define MAGPY_B200_ERR_NOMEM 4     /* request does not fit device or host memory        */
#define MAGPY_B200_ERR_COMM 5      /* NCCL / rendezvous error of the multi-GPU path      */

/* field shapes: same numbering as `enum field::options` (include/field.hpp:95-97)
 * and `cpdef enum options` (magpy/core.pyx:38-42). */
#define MAGPY_B200_FIELD_SINE 0
#define MAGPY_B200_FIELD_SQUARE 1
#define MAGPY_B200_FIELD_CONSTANT 2
#define MAGPY_B200_FIELD_SQUARE_FOURIER 3   /* discrete-orientation model only (field::square_fourier, lib/field.cpp:66-79) */

/* Gaussian transform applied to the Philox4x32-10 words (replaces lib/rng.cpp:14-24) */
#define MAGPY_B200_GAUSS_F32 0 /* Box-Muller in fp32 on the SFU (32-bit uniforms), widened to fp64 */
#define MAGPY_B200_GAUSS_F64 1 /* Box-Muller in fp64 from 53-bit uniforms                  */
#define MAGPY_B200_GAUSS_F32_PACKED 2 /* fp32 Box-Muller, 24-bit radius / 18-bit angle uniforms:
                                         one Philox block feeds two particle-steps          */

/* Statistics returned by every run (replaces the error codes the reference computes
 * and drops at lib/simulation.cpp:368-369 and the joblib progress lines of
 * magpy/model.py:204). */
typedef struct magpy_b200_stats {
    uint64_t steps_per_member;     /* integrator steps executed per member                 */
    uint64_t particle_steps;       /* members * particles * steps                          */
    uint64_t newton_iterations;    /* implicit only: total quasi-Newton iterations          */
    uint64_t newton_max_iterations;/* implicit only: largest count in a single step         */
    uint64_t newton_failures;      /* implicit only: steps that hit max_iter or a zero pivot */
    uint64_t kernel_launches;      /* CUDA kernels launched by this call                    */
    double device_ms;              /* CUDA-event time of the device work (launch stream)    */
    double integrate_ms;           /* CUDA-event time of the integration kernels alone      */
    uint64_t h2d_bytes;            /* host->device bytes copied by this call                */
    uint64_t d2h_bytes;            /* device->host bytes copied by this call                */
    uint64_t kernel_family;        /* MAGPY_B200_KERNEL_*: which integration kernel ran (ABI v3) */
    uint64_t kernel_variant;       /* heun_single: 1 = free register allocation (6 CTAs of 128 threads per SM), 7 = 7
                                      CTAs per SM (when the shard fits one such wave but not one of 6), 100 = latency
                                      variant for ensembles below one warp per SM sub-partition (field table
                                      prefetched a step pair ahead), 200 = persistent kernel over (time segment,
                                      member block) tasks for shards of more than one wave (no rounding up to whole
                                      warps per SM sub-partition), 300 = one integrator warp fed by three generator warps
                                      per 32 members (ensembles of at most 64 members per SM); 0 for the other kernels
                                      (ABI v4)                                                                  */
    /* host wall-clock of the blocking entry points (magpy_b200_simulate_ensemble[_multi]), ms (ABI v4): */
    double host_setup_ms;          /* plan creation: validation, schedule, allocations, uploads             */
    double host_run_ms;            /* launch + wait for the device                                          */
    double host_fetch_ms;          /* layout kernels + device->host copies into the caller's arrays         */
    double host_total_ms;          /* the whole call, including the release of the plan                     */
} magpy_b200_stats;

/* integration kernels (magpy_b200/csrc): reported in magpy_b200_stats.kernel_family */
#define MAGPY_B200_KERNEL_HEUN_SINGLE 1      /* heun_single.cu: one thread per single-particle member       */
#define MAGPY_B200_KERNEL_IMID_SINGLE 2      /* imid_single.cu                                              */
#define MAGPY_B200_KERNEL_HEUN_SMALL 3       /* small_heun.cu: one thread per cluster of 2..7 particles      */
#define MAGPY_B200_KERNEL_IMID_SMALL 4       /* small_imid.cu: 2..4 particles                                */
#define MAGPY_B200_KERNEL_HEUN_CLUSTER 5     /* cluster.cu: scalar dipolar sum from shared memory            */
#define MAGPY_B200_KERNEL_IMID_CLUSTER 6     /* cluster.cu                                                   */
#define MAGPY_B200_KERNEL_HEUN_CLUSTER_MMA 7 /* cluster_mma.cu: dipolar field as a matrix product (DMMA)     */
#define MAGPY_B200_KERNEL_IMID_SPLIT 8       /* small_imid.cu: one lane per particle (small ensembles, N = 2, 4) */
#define MAGPY_B200_KERNEL_HEUN_CLUSTER_BIG 10 /* cluster_big.cu: Heun beyond 128 particles, moments in global memory      */
#define MAGPY_B200_KERNEL_IMID_WARPS 12      /* small_imid.cu: one warp per particle (small ensembles, N = 2..4)  */
#define MAGPY_B200_KERNEL_IMID_CLUSTER_BIG 11 /* cluster_big.cu: implicit midpoint beyond 128 particles, iterates in global memory */
#define MAGPY_B200_KERNEL_IMID_CLUSTER_MMA 9 /* cluster_mma_imid.cu: implicit midpoint, dipolar field of every quasi-Newton
                                                iteration as a matrix product (DMMA), 8..128 particles             */

/* One ensemble of `n_members` independent clusters that share geometry and material
 * (radius, anisotropy, location, Ms, damping, T, field) and may differ in anisotropy
 * axes and initial magnetisation — replaces the joblib fan-out of
 * `EnsembleModel.simulate` (magpy/model.py:159-208) + one
 * `simulation::full_dynamics` call per member (lib/simulation.cpp:476-624). */
typedef struct magpy_b200_comm magpy_b200_comm; /* opaque: one rank of a one-process-per-GPU job (NCCL) */

typedef struct magpy_b200_ensemble {
    uint32_t abi_version;            /* = MAGPY_B200_ABI_VERSION                              */
    int32_t device;                  /* CUDA device ordinal                                   */
    uint64_t n_members;              /* R                                                     */
    uint32_t n_particles;            /* N                                                     */
    const double* radius;            /* [N]    m                                              */
    const double* anisotropy;        /* [N]    J/m^3                                          */
    const double* location;          /* [N][3] m                                              */
    const double* anisotropy_axis;   /* [N][3] if axis_stride==0 else [R][N][3]               */
    uint64_t axis_stride;            /* doubles between members: 0 (shared) or 3N             */
    const double* magnetisation_direction; /* [N][3] or [R][N][3] initial unit vectors        */
    uint64_t m0_stride;              /* 0 (shared) or 3N                                      */
    double magnetisation;            /* Ms, A/m                                               */
    double damping;                  /* alpha                                                 */
    double temperature;              /* K                                                     */
    int32_t renorm, interactions, use_implicit;
    double implicit_tol;             /* eps of the quasi-Newton solve                         */
    double time_step, end_time;      /* s                                                     */
    uint64_t max_samples;            /* S >= 2                                                */
    int32_t field_shape;             /* MAGPY_B200_FIELD_*                                    */
    double field_amplitude;          /* A/m                                                   */
    double field_frequency;          /* Hz                                                    */
    /* noise: Philox4x32-10 keyed by the member's seed, counter = (step, particle, member
     * index + stream_offset); or Wiener increments injected by the caller (RngArray
     * semantics, lib/rng.cpp:67-94) */
    const int64_t* seeds;            /* [R]                                                   */
    uint64_t stream_offset;          /* global index of member 0 (shards use disjoint ranges) */
    int32_t gauss_mode;              /* MAGPY_B200_GAUSS_*                                    */
    const double* injected_dw;       /* NULL, or [R][injected_steps][3N] unit-variance draws  */
    uint64_t injected_steps;
    /* outputs: host pointers, each may be NULL */
    double* out_time;                /* [S] s                                                 */
    double* out_field;               /* [S] A/m                                               */
    double* out_trajectories;        /* [R][N][3][S] A/m                                      */
    double* out_sums;                /* [S][4]: sum over members of cluster-summed Mx,My,Mz
                                        and of (cluster Mz)^2, A/m                            */
    double* out_final;               /* [R][N][3] state at the last sample, A/m               */
    /* Coarsened noise for convergence studies on common Brownian paths (test/convergence/task5.cpp:
     * 150-158: "coarse increments are sums of fine ones"): with L = noise_coarsen_log2 > 0 the unit-variance
     * increment of step s is 2^(-L/2) times the sum of the 2^L increments the packed Philox stream
     * (MAGPY_B200_GAUSS_F32_PACKED) assigns to the fine steps s 2^L ... (s+1) 2^L - 1, so runs with
     * time_step = dt 2^L, L = 0, 1, 2, ... see the same Wiener path.  Single-particle ensembles only. */
    uint32_t noise_coarsen_log2;
    /* Implicit midpoint only (ABI v3).  MAGPY_B200_NEWTON_REFERENCE (0, default): the reference's quasi-Newton
     * iteration, iterate by iterate (lib/integrators.cpp:576-651) — the parity mode.  MAGPY_B200_NEWTON_EXACT (1):
     * Newton's method with the exact Jacobian of the midpoint residual (hard part 1: "worthwhile for
     * throughput but must be opt-in"): the same implicit-midpoint equation solved to a tighter residual in ~3
     * iterations instead of ~20; trajectories differ from the reference's truncated iterates at the 1e-9 level per
     * step.  For clusters each particle's own exact Jacobian is used; the dipolar coupling between particles stays
     * out of the matrix (as in the reference), which leaves a fast linear convergence (4-5 iterations). */
    uint32_t implicit_newton;
    /* Per-member radii (ABI v3): 0 = `radius` holds the N radii shared by all members; N = `radius` holds
     * [R][N] values.  Supported for single-particle ensembles (N = 1) — a size distribution integrated in one
     * launch: with one particle the radius enters only through the thermal field strength sigma_i
     * (lib/simulation.cpp:531-535), which becomes a per-member array. */
    uint64_t radius_stride;
    /* Per-member temperatures (ABI v3): NULL, or [R] temperatures in K that replace `temperature` member by member.
     * Single-particle ensembles only, for the same reason (the temperature enters through sigma_i alone). */
    const double* member_temperature;
    /* ---- ABI v4 ---- */
    /* Global member indices (NULL, or [R]): word 1 of the Philox counter of member r is member_index[r] instead of
     * stream_offset + r, so that a member keeps its noise stream however the ensemble is cut into parameter groups,
     * shards and devices (values < 2^32). */
    const uint64_t* member_index;
    /* Per-member material parameters of single-particle ensembles (N = 1), each NULL or [R]; they replace
     * anisotropy[0] / damping / field_amplitude member by member — the reference's
     * `EnsembleModel(N, base, anisotropy=[...], damping=[...], field_amplitude=[...])` (magpy/model.py:146-156) in ONE
     * launch.  Anisotropy and damping change the member's reduced time scale (lib/simulation.cpp:514-520), so every
     * thread carries its own dt, sigma, field scale and zero-order-hold schedule (lib/simulation.cpp:342-355 evaluated per
     * member on the device in the same fp64 arithmetic).  Heun and implicit midpoint, field shapes constant and sine
     * (anisotropy / damping) or any shape (field amplitude); gauss_mode F32_PACKED or injected increments. */
    const double* member_anisotropy;
    const double* member_damping;
    const double* member_field_amplitude;   /* when set, out_field is the waveform for an amplitude of 1 A/m (member r saw
                                               member_field_amplitude[r] times it) */
    /* Multi-GPU (one process per GPU): when set, the [S][4] ensemble sums are all-reduced over the communicator
     * (ncclAllReduce, sum, fp64, in place on the device buffer, enqueued on the plan's stream after the last
     * integration kernel) before they are returned: out_sums then holds the sums over ALL ranks' members.  Every rank
     * of the communicator must make the matching call.  Replaces the gathering of pickled Results from the joblib
     * workers (magpy/model.py:204-207). */
    magpy_b200_comm* comm;
} magpy_b200_ensemble;
#define MAGPY_B200_NEWTON_REFERENCE 0
#define MAGPY_B200_NEWTON_EXACT 1

typedef struct magpy_b200_plan magpy_b200_plan; /* opaque: device-resident ensemble */

/* ---- library -------------------------------------------------------------------- */
int magpy_b200_abi_version(void);
const char* magpy_b200_last_error(void);
int magpy_b200_device_count(int* count);
/* Device buffers of finished calls stay cached in the device's stream-ordered memory pool so
 * that repeated calls do not pay cudaMalloc/cudaFree again; this returns them to the driver. */
int magpy_b200_release_cached_memory(int device);

/* Page-locked host buffers for OUTPUT arrays.  A device->host copy into pageable memory that has never been touched (a
 * fresh numpy.empty, a fresh std::vector) runs at ~5 GB/s on the B200 box: the driver stages it and every 4 KB page
 * faults on first write (profiles/r02_probe_e2e_small.log: 24 MB in 5.5 ms, 2.5 GB in 510 ms).  Into these buffers the
 * same copy is one DMA transfer at PCIe speed.  Freed blocks stay in a size-bucketed cache (at most
 * MAGPY_B200_PINNED_CACHE_MB, default 8192) so that repeated calls do not pay the pinning again;
 * magpy_b200_host_cache_release returns them to the system.  The Cython layer allocates the trajectory / final-state
 * arrays it returns from here (numpy arrays that own their block).  Replaces the element-wise copies into fresh numpy
 * arrays of magpy/core.pyx:186-203. */
int magpy_b200_host_alloc(size_t bytes, void** ptr);
int magpy_b200_host_free(void* ptr);
int magpy_b200_host_cache_release(void);

/* physical constants exactly as include/constants.hpp:10-12 (replaces
 * core.get_KB / get_mu0 / get_gamma, magpy/core.pyx:29-34) */
double magpy_b200_get_KB(void);
double magpy_b200_get_mu0(void);
double magpy_b200_get_gamma(void);

/* ---- one cluster ------------------------------------------------------------------
 * Replaces `simulation::full_dynamics` SI overload (include/simulation.hpp:74-94,
 * lib/simulation.cpp:476-624) as bound by magpy/core.pyx:72-93.  Argument meaning and
 * order follow that overload; outputs replace `std::vector<simulation::results>`
 * (include/simulation.hpp:32-48): out_time[S], out_field[S], out_m[N][3][S]. */
int magpy_b200_simulate(const double* radius, const double* anisotropy, const double* anisotropy_axis,
                        const double* magnetisation_direction, const double* location, size_t n_particles,
                        double magnetisation, double damping, double temperature, int renorm,
                        int interactions, int use_implicit, double eps, double time_step, double end_time,
                        size_t max_samples, int64_t seed, int field_shape, double field_amplitude,
                        double field_frequency, double* out_time, double* out_field, double* out_m,
                        magpy_b200_stats* stats);

/* ---- ensembles -------------------------------------------------------------------- */
/* Host buffers in, host buffers out (upload, integrate, reduce, download). */
int magpy_b200_simulate_ensemble(const magpy_b200_ensemble* args, magpy_b200_stats* stats);

/* The same over several GPUs of one box from ONE host thread: the members are cut into contiguous ranges of
 * ceil(R / n_devices) (member i keeps its seed and its Philox index whatever the device count, so results do not
 * depend on it), every device integrates its range concurrently on its own stream, per-member outputs land in the
 * caller's arrays at the member's global position and the [S][4] ensemble sums of the devices are added on the host
 * in device-list order (n_devices x 3 KB: no collective is needed inside one process).  `args->device` is ignored.
 * Replaces the `n_jobs` process pool of magpy/model.py:204-207. */
int magpy_b200_simulate_ensemble_multi(const magpy_b200_ensemble* args, const int* devices, int n_devices,
                                       magpy_b200_stats* stats);

/* Device-resident variant: create uploads inputs and allocates outputs once; run
 * re-integrates from the initial state (asynchronously on the plan's stream);
 * fetch copies the requested outputs into the host pointers of `args`. */
int magpy_b200_plan_create(const magpy_b200_ensemble* args, magpy_b200_plan** plan);
int magpy_b200_plan_run(magpy_b200_plan* plan);
int magpy_b200_plan_sync(magpy_b200_plan* plan, magpy_b200_stats* stats);
int magpy_b200_plan_fetch(magpy_b200_plan* plan, double* out_time, double* out_field,
                          double* out_trajectories, double* out_sums, double* out_final);
/* device address of the [S][4] ensemble-sum buffer (for an in-place NCCL all-reduce) */
int magpy_b200_plan_sums_device_ptr(magpy_b200_plan* plan, void** dptr, size_t* n_doubles);
int magpy_b200_plan_destroy(magpy_b200_plan* plan);

/* ---- multi-GPU: one process per GPU, one collective -----------------------------------------
 * The path shards by member index with no data-path exchange (magpy/model.py:204-207: independent members); the only
 * collective is ONE all-reduce (sum, fp64) of the [S][4] ensemble sums per pass, issued by the library through NCCL
 * (loaded with dlopen on first use; no link-time dependency).  A communicator is built either from an id the caller
 * distributes itself (magpy_b200_comm_unique_id on rank 0 -> any out-of-band channel -> magpy_b200_comm_create on every
 * rank: ncclGetUniqueId / ncclCommInitRank), or from the launcher's environment (RANK, WORLD_SIZE, LOCAL_RANK,
 * MASTER_ADDR, MASTER_PORT as set by torchrun): rank 0 serves the id over TCP on MASTER_PORT + 1
 * (MAGPY_B200_COMM_PORT overrides).  device < 0 = LOCAL_RANK. */
#define MAGPY_B200_COMM_ID_BYTES 128
#define MAGPY_B200_COMM_SUM 0
#define MAGPY_B200_COMM_MAX 1
int magpy_b200_comm_unique_id(uint8_t id[MAGPY_B200_COMM_ID_BYTES]);
int magpy_b200_comm_create(const uint8_t id[MAGPY_B200_COMM_ID_BYTES], int rank, int world_size, int device,
                           magpy_b200_comm** comm);
int magpy_b200_comm_create_from_env(int device, magpy_b200_comm** comm);
int magpy_b200_comm_rank(const magpy_b200_comm* comm, int* rank, int* world_size);
/* all-reduce of a small HOST array (staged through the device, ncclAllReduce): job-wide maxima of timings, ensemble
 * sums accumulated on the host over several parameter groups */
int magpy_b200_comm_allreduce(magpy_b200_comm* comm, double* host_values, size_t n, int op);
int magpy_b200_comm_barrier(magpy_b200_comm* comm);
int magpy_b200_comm_destroy(magpy_b200_comm* comm);

/* ---- host helpers that mirror reference arithmetic ---------------------------------- */
/* SI -> reduced units (lib/simulation.cpp:498-549): out[0]=V_av [1]=K_av [2]=H_k
 * [3]=time_factor [4]=dt_red [5]=T_red [6]=h0 [7]=f_red [8]=dipolar prefactor;
 * k_red/v_red/sigma are [N]. */
int magpy_b200_reduce_units(const double* radius, const double* anisotropy, size_t n_particles,
                            double magnetisation, double damping, double temperature, double time_step,
                            double end_time, double field_amplitude, double field_frequency,
                            double* k_red, double* v_red, double* sigma, double* out_scalars);
/* cumulative step count per sample of the zero-order-hold schedule
 * (lib/simulation.cpp:171-174, 342-355, 392-405): cum[k] = steps executed when sample k
 * is stored; sample k holds the state after cum[k]-1 steps. */
int magpy_b200_schedule(double dt_red, double t_end_red, size_t max_samples, uint64_t* cum);

/* ---- discrete-orientation model ------------------------------
 * Replaces `simulation::dom_ensemble_dynamics` (include/simulation.hpp:101-111, lib/simulation.cpp:660-766) as bound
 * by `simulate_dom` (magpy/core.pyx:205-278): the two-state master equation of a uniaxial particle in a field along
 * its axis (lib/dom.cpp:33-101), adaptive Cash-Karp RK45 (lib/integrators.cpp:152-251) with tolerance `time_step`,
 * first-order-hold samples.  Batched: `n_items` independent particles (volume, anisotropy, initial probabilities per
 * item; temperature, magnetisation, damping and the field shared), one GPU thread each — the reference call is
 * n_items = 1.  Outputs: out_time[S] (s), out_field[n_items][S] (A/m, i.e. the reduced field times the item's
 * H_k = 2K/(mu0 Ms) as magpy/core.pyx:224-225,264), out_mz[n_items][S] (p_0 - p_1, unitless as the reference returns
 * it), out_steps[n_items] accepted RK45 steps (may be NULL). */
int magpy_b200_simulate_dom(int device, size_t n_items, const double* volume, const double* anisotropy,
                            const double* initial_probabilities /* [n_items][2] */, double temperature,
                            double magnetisation, double damping, double time_step, double end_time, size_t max_samples,
                            int field_shape, double field_amplitude, double field_frequency, size_t field_n_components,
                            double* out_time, double* out_field, double* out_mz, uint64_t* out_steps);

/* ---- device self-tests used by the parity suite (tiny kernels) ----------------------- */
/* raw Philox4x32-10 words computed on the device */
int magpy_b200_philox_words(int device, const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* the n Gaussian draws member `member` / particle `particle` consumes at `step` (3 values) */
int magpy_b200_gaussians(int device, int64_t seed, uint64_t member, uint32_t particle, uint64_t first_step,
                         uint64_t n_steps, int gauss_mode, double* out /* [n_steps][3] */);
/* the 3x3 solve of one particle's quasi-Newton update as the implicit kernels run it (adjugate / Cramer's rule in
 * registers instead of dgesv's pivoted elimination, lib/optimisation.cpp:134): n systems A[n][9] (row major) x = b[n][3];
 * ok[i] = 0 where the determinant is exactly zero (dgesv's info > 0) */
int magpy_b200_solve3(int device, size_t n, const double* A, const double* b, double* x, int* ok);
/* Statistics of the Gaussian stream accumulated on the device (for tests at 1e9+ draws against lib/rng.cpp:14-24's
 * N(0,1)): n_members threads (members first_member ...) x n_steps steps x 3 draws, consumed exactly as the integration
 * kernels do.  hist[4096]: counts in bins of width 1/256 over [-8, 8); angle_hist[1024]: direction of the (x, y) pairs
 * that come from one Box-Muller pair, over one revolution (bin b = directions [b, b + 1) 2 pi / 1024 - pi 2^-18); moments[5] = {sum z, sum z^2, sum z^3, sum z^4, max |z|}. */
int magpy_b200_gaussian_stats(int device, int64_t seed, uint64_t first_member, uint64_t n_members, uint64_t n_steps,
                              int gauss_mode, uint64_t* hist, uint64_t* angle_hist, double* moments);
/* sustained FP64 FMA rate of the device (TFLOP/s), measured with a register-resident
 * DFMA chain kernel — the roofline denominator for the integration kernels */
int magpy_b200_fp64_peak(int device, double* tflops, double* sm_clock_mhz);
/* the same for the FP64 matrix instruction (DMMA.8x8x4, register-resident accumulator chains) — the roofline
 * denominator of the cluster kernel that evaluates the dipolar field as a matrix product (cluster_mma.cu) */
int magpy_b200_fp64_mma_peak(int device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* MAGPY_B200_H */
