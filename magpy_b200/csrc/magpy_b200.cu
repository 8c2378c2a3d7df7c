// magpy_b200.cu — host side of the B200-native ensemble sLLG integrator and its C ABI
// (include/magpy_b200.h).  Replaces, for one path only, the reference's
//   lib/simulation.cpp  (SI->reduced conversion :498-549, schedule :171-174/:342-405,
//                        rescale :610-621)
//   magpy/model.py:202-207 (per-member fan-out)
// and drives the kernels through the launchers of launch.h.  No CPU integration path exists here: without a
// CUDA device every compute entry point fails with MAGPY_B200_ERR_NO_DEVICE.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <unordered_map>
#include <new>
#include <string>
#include <vector>

#include "../../include/magpy_b200.h"
#include "common.cuh"
#include "launch.h"
#include "host_util.h"

namespace mbh {

thread_local std::string g_error;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}

int select_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(MAGPY_B200_ERR_NO_DEVICE,
                    "no CUDA device available (%s); magpy_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= n) return fail(MAGPY_B200_ERR_NO_DEVICE, "device %d out of range [0,%d)", device, n);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(MAGPY_B200_ERR_CUDA, "cudaSetDevice(%d) failed: %s", device, cudaGetErrorString(e));
    return MAGPY_B200_OK;
}

}  // namespace mbh

namespace {

// include/constants.hpp:10-12, digit for digit (MU0 is the reference's truncated value)
constexpr double kKB = 1.38064852e-23;
constexpr double kMU0 = 1.25663706e-6;
constexpr double kGYROMAG = 1.76086e11;

using mbh::fail;
using mbh::select_device;

#define CU_TRY(expr)                                                                                  \
    do {                                                                                              \
        cudaError_t e__ = (expr);                                                                     \
        if (e__ != cudaSuccess)                                                                       \
            return fail(e__ == cudaErrorMemoryAllocation ? MAGPY_B200_ERR_NOMEM : MAGPY_B200_ERR_CUDA, \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__);  \
    } while (0)

// ---- reference host arithmetic --------------------------------------------------------
struct Reduced {
    double V_av, K_av, H_k, tau, dt, T, h0, f, dip_pre;
    std::vector<double> k_red, v_red, sigma;
};

// lib/simulation.cpp:498-549 (same evaluation order)
void reduce_units(const double* radius, const double* anisotropy, size_t N, double Ms, double alpha, double T,
                  double dt, double t_end, double H0, double f, Reduced& out) {
    std::vector<double> vol(N);
    double vsum = 0.0, ksum = 0.0;
    for (size_t i = 0; i < N; ++i) {
        vol[i] = 4.0 / 3.0 * M_PI * radius[i] * radius[i] * radius[i];
        vsum += vol[i];
        ksum += anisotropy[i];
    }
    out.V_av = vsum / N;
    out.K_av = ksum / N;
    out.k_red.resize(N);
    out.v_red.resize(N);
    out.sigma.resize(N);
    for (size_t i = 0; i < N; ++i) {
        out.v_red[i] = vol[i] / out.V_av;
        out.k_red[i] = anisotropy[i] / out.K_av;
    }
    out.H_k = 2 * out.K_av / kMU0 / Ms;
    out.tau = kGYROMAG * kMU0 * out.H_k / (1 + alpha * alpha);
    out.dt = dt * out.tau;
    out.T = t_end * out.tau;
    for (size_t i = 0; i < N; ++i)
        out.sigma[i] = std::sqrt(alpha * kKB * T / (out.K_av * vol[i]) / (1 + alpha * alpha));
    out.h0 = H0 / out.H_k;
    out.f = f / out.tau;
    out.dip_pre = kMU0 * Ms * Ms / 8.0 / M_PI / out.K_av;  // lib/field.cpp:212-215
}

// lib/field.cpp:23-54 + lib/simulation.cpp:553-573
double applied_field(int shape, double t, double h, double f) {
    switch (shape) {
        case MAGPY_B200_FIELD_SINE: return h * std::sin(2 * M_PI * f * t);
        case MAGPY_B200_FIELD_SQUARE: return h * (int(t * f * 2) % 2 ? -1 : 1);
        default: return h;
    }
}

// Zero-order-hold schedule (lib/simulation.cpp:342-355) without walking every step:
// cum[k] = smallest step count s >= cum[k-1] with fl(s*dt) > fl(k*Ts).
int build_schedule(double dt, double T, size_t S, std::vector<uint64_t>& cum) {
    cum.assign(S, 0);
    const double Ts = T / (S - 1);
    uint64_t prev = 0;
    for (size_t k = 1; k < S; ++k) {
        const double lim = (double)(unsigned int)k * Ts;
        double est = std::floor(lim / dt);
        if (!(est >= 0.0) || est > 4.0e9) return 1;
        uint64_t s = (uint64_t)est;
        s = s > 2 ? s - 2 : 0;
        if (s < prev) s = prev;
        while ((double)s * dt <= lim) ++s;          // first s that breaks the loop condition
        while (s > prev && (double)(s - 1) * dt > lim) --s;
        if (s > 0xFFFFFFF0ull) return 1;            // the reference's step counter is 32-bit
        cum[k] = s;
        prev = s;
    }
    return 0;
}

// ---- device buffers -------------------------------------------------------------------
// Device allocations of a plan come from the device's stream-ordered memory pool with the
// release threshold lifted, so the second and later calls of a process reuse the blocks of the
// first instead of paying cudaMalloc/cudaFree (tens of ms per ensemble call) again.
int enable_pool(int device) {
    static bool done[64] = {};
    if (device < 0 || device >= 64 || done[device]) return MAGPY_B200_OK;
    cudaMemPool_t pool;
    CU_TRY(cudaDeviceGetDefaultMemPool(&pool, device));
    unsigned long long keep = ~0ull;
    CU_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    done[device] = true;
    return MAGPY_B200_OK;
}

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaStream_t pool_stream = nullptr;   // non-null: stream-ordered pool allocation
    cudaError_t alloc(size_t count, cudaStream_t stream = nullptr) {
        release();
        n = count;
        pool_stream = stream;
        if (count == 0) return cudaSuccess;
        return stream ? cudaMallocAsync(&p, count * sizeof(T), stream) : cudaMalloc(&p, count * sizeof(T));
    }
    void release() {
        if (p) {
            if (pool_stream) cudaFreeAsync(p, pool_stream);
            else cudaFree(p);
        }
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
};

struct Chunk {
    uint64_t j0, j1;
    uint32_t k0, k1;
};

// per-member schedules (MP kernels) differ from the shared one by a step at most (exact ties of k Ts against s dt):
// the field table of a launch covers this many extra steps on either side
constexpr uint64_t kMpMargin = 4;

// K1b: at most this many time segments per launch, each at least kBalMinSteps long
constexpr uint32_t kBalMaxSegments = 128;   // 125,000 members x 1e5 steps: 64 segments 54.9 ms, 128 54.0, 256 54.5
constexpr uint64_t kBalMinSteps = 512;

}  // namespace

struct magpy_b200_plan {
    int device = 0;
    uint64_t R = 0;
    uint32_t N = 0, n = 0;
    uint64_t S = 0;
    bool implicit = false, want_traj = false, injected = false;
    int gauss_mode = 0, field_shape = MAGPY_B200_FIELD_CONSTANT;
    uint32_t coarsen = 0;   // noise_coarsen_log2
    double Ms = 0;
    Reduced red;
    std::vector<uint64_t> target;   // state index stored by sample k
    std::vector<Chunk> chunks;
    uint64_t total_steps = 0;
    mb::RunParams base{};
    unsigned grid = 0;
    dim3 block{1, 1, 1};
    size_t smem = 0;
    int np = 1;
    int layout = 0;        // Heun cluster kernel: see cluster.cu
    bool use_table = false;
    bool axis_z = false;   // N = 1 and one shared easy axis exactly along +z: specialised Heun kernel
    bool k1_balanced = false;   // K1b: multi-wave shards run as a persistent kernel over (time segment, member block) tasks
    unsigned bal_phys = 0;      //   ... its physical grid (resident CTAs)
    std::vector<uint32_t> bal_segments;   //   ... segments per chunk (1 = plain launch)
    bool k1_split = false;      // K1s: ensembles of at most 64 members per SM: integrator warp + three generator warps per 32 members
    int k1_min_blocks = 1; // K1: register-allocation variant (resident CTAs per SM asked of ptxas), see choose_k1_variant
    bool small = false;    // few particles: one thread per cluster, all moments in registers
    bool split = false;    //   ... implicit, one lane per particle (imid_split_kernel)
    bool warps = false;    //   ... implicit, one warp per particle (imid_warps_kernel)
    bool mma = false;      // Heun cluster kernel on the FP64 MMA path (cluster_mma.cu)
    bool imid_mma = false; // implicit cluster kernel on the FP64 MMA path (cluster_mma_imid.cu)
    bool big = false;      // beyond 128 particles: moments in global memory (cluster_big.cu), Heun and implicit midpoint
    bool one_buf = false;  //   ... with one shared-memory moment buffer
    uint32_t G = 0;        //   ... particle groups of 8
    uint32_t mma_full = 0, mma_tail = 0;   //   ... member distribution over CTAs (see choose_mma)
    bool mma_dglobal = false;              //   ... matrix read from global memory (N > 64)
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    std::vector<cudaEvent_t> ev_k;   // pairs around each integration launch
    DevBuf<double> d_state_t, d_state_u;
    DevBuf<double> d_state0, d_state, d_axis, d_kred, d_sig, d_dip, d_dmat, d_vred, d_traj, d_sums, d_partial, d_tab, d_dW, d_stage;
    DevBuf<int64_t> d_seeds;
    DevBuf<uint32_t> d_member_idx, d_member_j, d_bal_seg_k;
    DevBuf<unsigned int> d_bal_sync;   // [1 + grid]: task counter, then one flag per member block
    DevBuf<double> d_mp;            // per-member material parameters: alpha | dt | h0 | Ts, [4][R]
    bool mp = false;                // per-member anisotropy / damping / field amplitude (N = 1): MP kernel instantiations
    magpy_b200_comm* comm = nullptr;   // all-reduce the sums over this communicator at the end of every run
    DevBuf<uint64_t> d_target;
    DevBuf<unsigned long long> d_newton;
    uint64_t launches = 0, h2d = 0, d2h = 0;
    uint64_t max_chunk_steps = 0;
    uint32_t max_chunk_samples = 0;
    bool ran = false;
    bool unit_field = false;   // per-member field amplitudes: out_field is the waveform for 1 A/m

    ~magpy_b200_plan() {
        cudaSetDevice(device);
        d_state_t.release(); d_state_u.release(); d_state0.release(); d_state.release(); d_axis.release(); d_kred.release(); d_sig.release(); d_dip.release();
        d_dmat.release(); d_vred.release();
        d_traj.release(); d_sums.release(); d_partial.release(); d_tab.release(); d_dW.release(); d_stage.release();
        d_seeds.release(); d_member_idx.release(); d_member_j.release(); d_bal_seg_k.release(); d_bal_sync.release(); d_mp.release(); d_target.release(); d_newton.release();
        if (stream) cudaStreamSynchronize(stream);
        for (auto e : ev_k) cudaEventDestroy(e);
        if (ev_begin) cudaEventDestroy(ev_begin);
        if (ev_end) cudaEventDestroy(ev_end);
        if (stream) cudaStreamDestroy(stream);
    }
};

namespace {

#define LAUNCH_TRY(expr)                                                                             \
    do {                                                                                             \
        cudaError_t e__ = (expr);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return fail(MAGPY_B200_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
        pl->launches++;                                                                              \
    } while (0)

int plan_noise(const magpy_b200_plan* pl) {
    if (pl->injected) return mb::NOISE_INJECTED;
    if (pl->coarsen > 0) return mb::NOISE_PHILOX_COARSE;
    if (pl->gauss_mode == MAGPY_B200_GAUSS_F64) return mb::NOISE_PHILOX_F64;
    if (pl->gauss_mode == MAGPY_B200_GAUSS_F32) return mb::NOISE_PHILOX_F32;
    return mb::NOISE_PHILOX_PACKED;
}

// ---- kernel dispatch ------------------------------------------------------------------
int launch_integrate(magpy_b200_plan* pl, const mb::RunParams& P) {
    const int noise = plan_noise(pl);
    const bool tab = pl->use_table;
    if (pl->N == 1) {
        if (pl->implicit) LAUNCH_TRY(mb::launch_imid_single(noise, tab, pl->axis_z, pl->grid, pl->stream, P));
        else if (pl->k1_split) LAUNCH_TRY(mb::launch_heun_single_split(tab, pl->axis_z, (int)pl->block.x / 32 - 1, pl->grid, pl->stream, P));
        else LAUNCH_TRY(mb::launch_heun_single(noise, tab, pl->axis_z, pl->k1_min_blocks, pl->grid, pl->stream, P));
    } else if (pl->small) {
        if (pl->warps) LAUNCH_TRY(mb::launch_imid_warps(noise, tab, pl->N, pl->grid, pl->stream, P));
        else if (pl->split) LAUNCH_TRY(mb::launch_imid_split(noise, tab, pl->N, pl->grid, pl->stream, P));
        else if (pl->implicit) LAUNCH_TRY(mb::launch_imid_small(noise, tab, pl->N, pl->grid, pl->stream, P));
        else LAUNCH_TRY(mb::launch_heun_small(noise, tab, pl->N, pl->grid, pl->stream, P));
    } else if (pl->mma) {
        LAUNCH_TRY(mb::launch_heun_cluster_mma(noise, tab, pl->one_buf, pl->mma_full, pl->grid - pl->mma_full, pl->block.x, pl->smem,
                                               pl->stream, P));
        if (pl->mma_full > 0 && pl->grid > pl->mma_full) pl->launches++;   // whole waves + partial last wave
    } else if (pl->big) {
        if (pl->implicit) LAUNCH_TRY(mb::launch_imid_cluster_big(noise, tab, pl->grid, pl->stream, P));
        else LAUNCH_TRY(mb::launch_heun_cluster_big(noise, tab, pl->grid, pl->stream, P));
    } else if (pl->imid_mma) {
        LAUNCH_TRY(mb::launch_imid_cluster_mma(noise, tab, pl->grid, pl->block.x, pl->smem, pl->stream, P));
    } else if (pl->implicit) {
        LAUNCH_TRY(mb::launch_imid_cluster(noise, tab, pl->np, dim3(pl->grid), pl->block, pl->smem, pl->stream, P));
    } else {
        LAUNCH_TRY(mb::launch_heun_cluster(noise, tab, pl->np, pl->layout, dim3(pl->grid), pl->block, pl->smem, pl->stream, P));
    }
    return MAGPY_B200_OK;
}

int launch_transpose(magpy_b200_plan* pl, const double* in, double* out, uint64_t batches, uint64_t rows,
                     uint64_t cols, uint64_t in_bs, uint64_t in_rs, uint64_t out_bs, uint64_t out_rs,
                     double scale) {
    // batches go through gridDim.z and row tiles through gridDim.y, both in slices of 65535
    const uint64_t row_slice = 65535ull * 32;
    for (uint64_t b0 = 0; b0 < batches; b0 += 65535) {
        const uint64_t nb = std::min<uint64_t>(65535, batches - b0);
        for (uint64_t r0 = 0; r0 < rows; r0 += row_slice) {
            const uint64_t nr = std::min<uint64_t>(row_slice, rows - r0);
            // element (b, row, col): in[b in_bs + row in_rs + col] -> out[b out_bs + col out_rs + row]
            LAUNCH_TRY(mb::launch_transpose(in + b0 * in_bs + r0 * in_rs, out + b0 * out_bs + r0, nr, cols, nb, in_bs, in_rs,
                                            out_bs, out_rs, scale, pl->stream));
        }
    }
    return MAGPY_B200_OK;
}

int validate(const magpy_b200_ensemble* a) {
    if (!a) return fail(MAGPY_B200_ERR_BAD_ARG, "args is NULL");
    if (a->abi_version != MAGPY_B200_ABI_VERSION)
        return fail(MAGPY_B200_ERR_BAD_ARG, "abi_version %u != %d", a->abi_version, MAGPY_B200_ABI_VERSION);
    if (a->n_members == 0) return fail(MAGPY_B200_ERR_BAD_ARG, "n_members must be >= 1");
    if (a->n_members > 0xFFFFFFFFull) return fail(MAGPY_B200_ERR_BAD_ARG, "n_members must be < 2^32");
    if (a->n_particles == 0) return fail(MAGPY_B200_ERR_BAD_ARG, "n_particles must be >= 1");
    if (!a->radius || !a->anisotropy || !a->location || !a->anisotropy_axis || !a->magnetisation_direction)
        return fail(MAGPY_B200_ERR_BAD_ARG, "radius/anisotropy/location/anisotropy_axis/magnetisation_direction must not be NULL");
    const uint64_t n = 3ull * a->n_particles;
    if (a->axis_stride != 0 && a->axis_stride != n) return fail(MAGPY_B200_ERR_BAD_ARG, "axis_stride must be 0 or 3N");
    if (a->m0_stride != 0 && a->m0_stride != n) return fail(MAGPY_B200_ERR_BAD_ARG, "m0_stride must be 0 or 3N");
    if (a->radius_stride != 0 && (a->radius_stride != a->n_particles || a->n_particles != 1))
        return fail(MAGPY_B200_ERR_BAD_ARG, "per-member radii (radius_stride = N) are supported for single-particle ensembles only");
    if (a->member_temperature && a->n_particles != 1)
        return fail(MAGPY_B200_ERR_BAD_ARG, "per-member temperatures are supported for single-particle ensembles only");
    if (a->max_samples < 2) return fail(MAGPY_B200_ERR_BAD_ARG, "max_samples must be >= 2 (lib/simulation.cpp:174)");
    if (a->max_samples > 0x7FFFFFFFull) return fail(MAGPY_B200_ERR_BAD_ARG, "max_samples too large");
    if (!(a->time_step > 0.0) || !(a->end_time > 0.0)) return fail(MAGPY_B200_ERR_BAD_ARG, "time_step and end_time must be > 0");
    if (a->field_shape != MAGPY_B200_FIELD_SINE && a->field_shape != MAGPY_B200_FIELD_SQUARE &&
        a->field_shape != MAGPY_B200_FIELD_CONSTANT)
        return fail(MAGPY_B200_ERR_BAD_ARG, "Must specify valid field::options enum (lib/simulation.cpp:570-572)");
    if (!a->injected_dw && !a->seeds) return fail(MAGPY_B200_ERR_BAD_ARG, "seeds must not be NULL unless injected_dw is given");
    if (a->gauss_mode != MAGPY_B200_GAUSS_F32 && a->gauss_mode != MAGPY_B200_GAUSS_F64 &&
        a->gauss_mode != MAGPY_B200_GAUSS_F32_PACKED)
        return fail(MAGPY_B200_ERR_BAD_ARG, "gauss_mode must be MAGPY_B200_GAUSS_F32, _F64 or _F32_PACKED");
    if (!(a->magnetisation > 0.0)) return fail(MAGPY_B200_ERR_BAD_ARG, "magnetisation must be > 0");
    if (a->noise_coarsen_log2 > 0) {
        if (a->n_particles != 1 || a->injected_dw || a->gauss_mode != MAGPY_B200_GAUSS_F32_PACKED)
            return fail(MAGPY_B200_ERR_BAD_ARG, "noise_coarsen_log2 needs a single-particle ensemble with the packed Philox stream");
        if (a->noise_coarsen_log2 > 20) return fail(MAGPY_B200_ERR_BAD_ARG, "noise_coarsen_log2 must be <= 20");
    }
    if (a->implicit_newton != MAGPY_B200_NEWTON_REFERENCE && a->implicit_newton != MAGPY_B200_NEWTON_EXACT)
        return fail(MAGPY_B200_ERR_BAD_ARG, "implicit_newton must be MAGPY_B200_NEWTON_REFERENCE or MAGPY_B200_NEWTON_EXACT");
    if (a->member_anisotropy || a->member_damping || a->member_field_amplitude) {
        if (a->n_particles != 1)
            return fail(MAGPY_B200_ERR_BAD_ARG, "per-member anisotropy / damping / field amplitude are supported for single-particle ensembles only");
        if (!a->injected_dw && a->gauss_mode != MAGPY_B200_GAUSS_F32_PACKED)
            return fail(MAGPY_B200_ERR_BAD_ARG, "per-member anisotropy / damping / field amplitude need gauss_mode F32_PACKED (or injected increments)");
        if (a->noise_coarsen_log2 > 0)
            return fail(MAGPY_B200_ERR_BAD_ARG, "per-member material parameters cannot be combined with noise_coarsen_log2");
        if ((a->member_anisotropy || a->member_damping) && a->field_shape == MAGPY_B200_FIELD_SQUARE)
            return fail(MAGPY_B200_ERR_BAD_ARG, "per-member anisotropy / damping change the member's time scale: the square "
                        "wave's switching instants would need per-member evaluation (use one ensemble per value)");
    }
    if (a->member_index)
        for (uint64_t r = 0; r < a->n_members; ++r)
            if (a->member_index[r] > 0xFFFFFFFFull) return fail(MAGPY_B200_ERR_BAD_ARG, "member_index[%llu] must be < 2^32", (unsigned long long)r);
    if (a->comm && mbh::comm_device(a->comm) != a->device)
        return fail(MAGPY_B200_ERR_BAD_ARG, "the communicator was created for device %d, the ensemble runs on device %d",
                    mbh::comm_device(a->comm), a->device);
    if (a->n_particles > 2048)   // the reference has no limit in its code; its dense (3N)^3 work arrays set one long before this
        return fail(MAGPY_B200_ERR_BAD_ARG, "at most 2048 particles per cluster");
    return MAGPY_B200_OK;
}

// K1 (heun_single.cu) is bound by the issue time of a warp-step (~145 cycles per sub-partition): a wave with k
// CTAs per SM takes k issue times per step, and a lone warp is hardly faster than that (155 cycles).  With ptxas left
// alone the kernel takes 76 registers = 6 resident CTAs per SM (888 per device).  Measured over shard sizes
// (profiles/r02_probe_k1_variants.log): a shard slightly larger than one wave does NOT pay a whole extra wave — its few
// tail CTAs run one per SM (1M members over 8 GPUs = 977 CTAs: 92 % of the large-ensemble rate) — but as the tail grows
// towards one CTA on every SM the block scheduler doubles some SMs up (1036 CTAs: 86 %).  The 7-CTA instantiation
// (66 registers, no spills) holds up to 1036 CTAs in ONE wave at 97 %; an 8-CTA one (62 registers) was measured too and
// never beats two waves of 6, so it is not built.  MAGPY_B200_K1_MIN_BLOCKS=1|7 overrides.
void choose_k1_variant(magpy_b200_plan* pl, bool renorm, bool mp) {
    pl->k1_min_blocks = 1;
    if (pl->N != 1 || pl->implicit || plan_noise(pl) != mb::NOISE_PHILOX_PACKED) return;
    if (const char* env = std::getenv("MAGPY_B200_K1_MIN_BLOCKS")) {
        const int v = std::atoi(env);
        if (v == 1 || v == 7 || v == mb::K1_LATENCY) { pl->k1_min_blocks = v; return; }
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pl->device);
    // K1s (heun_single_split.cu): up to 64 members per SM leave at least two of an SM's four sub-partitions idle; the step is
    // split over an integrator warp and three generator warps (BASELINE config 1, 1000 members x 1e5 steps: 9.52 -> 7.44 ms;
    // in a sine field 10.63 -> 7.76 ms; 9472 members 9.49 -> 7.80 ms; every shape of axis / renorm / field gains, 1-50 %:
    // profiles/r02_probe_c1_split_v2.log, r02_probe_k1s_shapes.log, r02_probe_k1s_producers.log).  Not for per-member material parameters.
    // MAGPY_B200_K1_SPLIT=0|1 overrides.
    {
        const uint64_t split_grid = (pl->R + 31) / 32;
        bool split = !mp && split_grid <= 2 * (uint64_t)sms;
        if (const char* env = std::getenv("MAGPY_B200_K1_SPLIT")) split = !mp && std::atoi(env) != 0 && split_grid < 0x7fffffffull;
        if (split) {
            pl->k1_split = true;
            pl->grid = (unsigned)split_grid;
            // one CTA per SM: three generator warps on the SM's idle sub-partitions; two CTAs per SM: one generator each, so
            // that every warp still has a sub-partition to itself (MAGPY_B200_K1_SPLIT_PRODUCERS=1|3 overrides)
            int producers = split_grid <= (uint64_t)sms ? 3 : 1;
            if (const char* env = std::getenv("MAGPY_B200_K1_SPLIT_PRODUCERS")) producers = std::atoi(env) == 1 ? 1 : 3;
            pl->block = dim3(32 * (1 + producers));
            return;
        }
    }
    // below ~4 warps per SM sub-partition a step is (mostly) one warp's own in-order stream — 11 dependent levels of FP64
    // instructions at 8.1 cycles of latency plus the serial issue of each level's members: 187 cycles per step however small
    // the ensemble — and an applied-field table entry fetched at its point of use adds its L2 latency to every step (334
    // cycles): the latency variant fetches it a pair ahead (profiles/r02_probe_c1.log: 1000 members 17.0 -> 11.3 ms per 1e5
    // steps, 37,888 members 24.5 -> 17.6 ms, break-even near 76k members = 4 CTAs per SM)
    if (pl->use_table && pl->grid <= (unsigned)(4 * sms)) { pl->k1_min_blocks = mb::K1_LATENCY; return; }
    const int free_ctas = mb::heun_single_resident_ctas(pl->use_table, pl->axis_z, renorm, 1);
    const int tight_ctas = mb::heun_single_resident_ctas(pl->use_table, pl->axis_z, renorm, 7);
    if (free_ctas > 0 && tight_ctas > free_ctas && pl->grid > (unsigned)(sms * free_ctas) &&
        pl->grid <= (unsigned)(sms * tight_ctas))
        pl->k1_min_blocks = 7;
}

// K2m (cluster_mma.cu) applies to Heun clusters of 8..128 interacting particles (above 64 the packed matrix stays in
// global memory: measured 2.0-2.4x the scalar kernel, profiles/r01_probe_cluster_mma_global.log).  Particles are handled in groups of
// 8 (rows of the 8x8x4 MMA), so a cluster that fills its last group badly does padded work.  Measured
// (profiles/r01_probe_cluster_mma_fill.log): the matrix kernel wins when N^2 / (8 G)^2 >= 0.7 and for every N > 32;
// the scalar kernel keeps the rest unless MAGPY_B200_CLUSTER_KERNEL=mma asks otherwise (=simt forces the scalar
// kernel everywhere).
bool choose_mma(const magpy_b200_ensemble* a, magpy_b200_plan* pl) {
    const uint32_t N = pl->N;
    if (pl->implicit || N < 8 || N > 128 || a->interactions == 0) return false;
    const char* force = std::getenv("MAGPY_B200_CLUSTER_KERNEL");
    if (force && std::strcmp(force, "simt") == 0) return false;
    const uint32_t G = (N + 7) / 8;
    const double fill = (double)N * N / (64.0 * G * G);
    if (!(force && std::strcmp(force, "mma") == 0) && fill < 0.7 && N <= 32) return false;
    pl->mma_dglobal = N > 64;   // the packed matrix no longer fits in shared memory: read it from global memory
    uint32_t MH = std::min<uint32_t>(8, 16 / G);
    const size_t cap = 227 * 1024;
    for (; MH >= 1; --MH) {
        const size_t MB = 16 * MH, LD = MB + 4;
        const size_t dmat = pl->mma_dglobal ? 0 : (size_t)G * (G + 1) / 2 * 576 * 8, mom = (size_t)24 * G * LD * 8,
                     red = (size_t)G * 3 * MB * 8;
        if (dmat + 2 * mom + red <= cap) { pl->one_buf = false; pl->smem = dmat + 2 * mom + red; break; }
        if (dmat + mom + red <= cap) { pl->one_buf = true; pl->smem = dmat + mom + red; break; }
    }
    if (MH == 0) return false;
    pl->mma = true;
    pl->G = G;
    pl->np = 1;
    pl->block = dim3(32 * G * MH);
    // whole waves of full CTAs (one CTA per SM: shared memory) + a last wave spread over all SMs in column tiles of 8
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pl->device);
    const uint64_t MB = 16 * MH, full_ctas = pl->R / MB;
    const uint64_t threads = 32ull * G * MH;
    const uint64_t per_sm = std::max<uint64_t>(1, std::min<uint64_t>({cap / pl->smem, 65536 / (128 * threads), 2048 / threads}));
    const uint64_t ctas_per_wave = (uint64_t)sms * per_sm;
    pl->mma_full = (uint32_t)(full_ctas / ctas_per_wave * ctas_per_wave);
    // few particle groups: a CTA's time is set by the update / noise latency of a warp, not by its column tiles —
    // thinning the CTAs of the last wave buys nothing there (measured, N = 8), so it stays a plain ragged grid
    if (G < 2) pl->mma_full = (uint32_t)((pl->R + MB - 1) / MB);
    const uint64_t rest = pl->R > (uint64_t)pl->mma_full * MB ? pl->R - (uint64_t)pl->mma_full * MB : 0;
    pl->mma_tail = (uint32_t)MB;
    uint64_t tail_ctas = 0;
    if (rest > 0) {
        const uint64_t tiles = (rest + 7) / 8, per_cta = std::min<uint64_t>(MB / 8, (tiles + ctas_per_wave - 1) / ctas_per_wave);
        pl->mma_tail = (uint32_t)(8 * per_cta);
        tail_ctas = (rest + pl->mma_tail - 1) / pl->mma_tail;
    }
    pl->grid = (unsigned)(pl->mma_full + tail_ctas);
    return true;
}

// K4m (cluster_mma_imid.cu): implicit midpoint for clusters of 8..128 particles.  One warp per (particle group of 8, column
// tile of 8 members): G * CT <= 16 warps; the packed matrix in shared memory up to 64 particles, in global memory / L2
// above; ONE moment buffer.  MAGPY_B200_CLUSTER_KERNEL=simt keeps the scalar kernel (cluster.cu, <= 64 particles).
bool choose_imid_mma(const magpy_b200_ensemble* a, magpy_b200_plan* pl) {
    const uint32_t N = pl->N;
    if (!pl->implicit || N < 8 || N > 128) return false;
    const char* force = std::getenv("MAGPY_B200_CLUSTER_KERNEL");
    if (force && std::strcmp(force, "simt") == 0 && N <= 40) return false;
    // measured (profiles/r02_probe_imid_cluster.log): 8 particles 1.67e9 vs 1.96e9 particle-steps/s for the scalar kernel,
    // 16: 1.35e9 vs 1.33e9, 32: 1.03e9 vs 0.89e9; 64 and more only fit here
    if (!(force && std::strcmp(force, "mma") == 0) && N <= 16) return false;
    (void)a;
    const uint32_t G = (N + 7) / 8;
    pl->mma_dglobal = N > 64;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pl->device);
    // column tiles per CTA: as many as 16 warps allow, fewer when the ensemble is too small to give every SM a CTA
    uint32_t CT = std::min<uint32_t>(15, std::max<uint32_t>(1, 16 / G));   // one named barrier (1..15) per column tile
    const uint64_t tiles = (pl->R + 7) / 8;
    CT = (uint32_t)std::min<uint64_t>(CT, std::max<uint64_t>(1, tiles / (uint64_t)sms));
    const size_t cap = 227 * 1024;
    for (; CT >= 1; --CT) {
        const size_t MB = 8 * CT, LD = MB + 4;
        const size_t dmat = pl->mma_dglobal ? 0 : (size_t)G * (G + 1) / 2 * 576 * 8;
        pl->smem = dmat + (size_t)24 * G * LD * 8 + (size_t)3 * G * MB * 8;
        if (pl->smem <= cap) break;
    }
    if (CT == 0) return false;
    pl->imid_mma = true;
    pl->G = G;
    pl->np = 1;
    pl->block = dim3(32 * G * CT);
    pl->grid = (unsigned)((pl->R + 8 * CT - 1) / (8 * CT));
    return true;
}

int plan_build(const magpy_b200_ensemble* a, magpy_b200_plan* pl) {
    int rc = validate(a);
    if (rc) return rc;
    rc = select_device(a->device);
    if (rc) return rc;
    pl->device = a->device;
    pl->R = a->n_members;
    pl->N = a->n_particles;
    pl->n = 3 * pl->N;
    pl->S = a->max_samples;
    pl->implicit = a->use_implicit != 0;
    pl->gauss_mode = a->gauss_mode;
    pl->coarsen = a->noise_coarsen_log2;
    pl->field_shape = a->field_shape;
    pl->Ms = a->magnetisation;
    pl->injected = a->injected_dw != nullptr;
    pl->want_traj = a->out_trajectories != nullptr;
    const uint64_t R = pl->R, n = pl->n;
    const uint32_t N = pl->N;

    reduce_units(a->radius, a->anisotropy, N, a->magnetisation, a->damping, a->temperature, a->time_step,
                 a->end_time, a->field_amplitude, a->field_frequency, pl->red);
    const Reduced& rd = pl->red;
    if (!(rd.dt > 0.0) || !std::isfinite(rd.dt) || !std::isfinite(rd.T))
        return fail(MAGPY_B200_ERR_BAD_ARG, "non-finite reduced time step (check anisotropy / magnetisation)");

    std::vector<uint64_t> cum;
    if (build_schedule(rd.dt, rd.T, pl->S, cum))
        return fail(MAGPY_B200_ERR_BAD_ARG, "end_time/time_step exceeds the 32-bit step counter of the reference");
    pl->target.resize(pl->S);
    pl->target[0] = 0;
    for (size_t k = 1; k < pl->S; ++k) pl->target[k] = cum[k] - 1;
    pl->total_steps = pl->target[pl->S - 1];
    if (pl->injected && a->injected_steps < pl->total_steps)
        return fail(MAGPY_B200_ERR_BAD_ARG, "injected_dw holds %llu steps but the schedule needs %llu",
                    (unsigned long long)a->injected_steps, (unsigned long long)pl->total_steps);

    // launch geometry
    if (N == 1) {
        pl->block = dim3(mb::SINGLE_THREADS);
        pl->grid = (unsigned)((R + mb::SINGLE_THREADS - 1) / mb::SINGLE_THREADS);
        pl->smem = 0;
        pl->np = 1;
    } else if (N <= 4 || (!pl->implicit && N <= 7)) {   // one thread per cluster (small_heun.cu: 2..7, small_imid.cu: 2..4)
        pl->small = true;
        pl->block = dim3(mb::SINGLE_THREADS);
        pl->grid = (unsigned)((R + mb::SINGLE_THREADS - 1) / mb::SINGLE_THREADS);
        pl->smem = 0;
        pl->np = 1;
        // implicit tetramers that cannot give every warp scheduler two warps at one thread per cluster are latency
        // bound: one lane per particle instead (small_imid.cu, imid_split_kernel; a dimer's two particles already
        // interleave in one thread, so it gains nothing).
        // MAGPY_B200_SMALL_KERNEL=split|thread overrides the choice.
        if (pl->implicit && N <= 4) {
            int sms = 148;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pl->device);
            bool split = N == 4 && R < (uint64_t)sms * 4 * 2 * 32;   // measured (profiles/r01_probe_c2.log): a gain for N = 4 only
            // one WARP per particle (imid_warps_kernel) for trimers / tetramers of at most one CTA per SM: 1000 tetramers 13.4 ms
            // against 16.6 (one lane per particle) and 28.1 (one thread per cluster); dimers gain 4-7 % with general easy axes
            // and LOSE with aligned ones (config 2: 14.5 against 11.2 ms), so they stay on one thread per cluster
            // (profiles/r02_probe_c2.log)
            bool warps = N >= 3 && R <= (uint64_t)sms * 32;
            if (warps) split = false;
            if (const char* force = std::getenv("MAGPY_B200_SMALL_KERNEL")) {
                if (std::strcmp(force, "split") == 0) split = (N == 2 || N == 4) && R * N < 0xFFFFFFFFull;
                else if (std::strcmp(force, "thread") == 0) { split = false; warps = false; }
                else if (std::strcmp(force, "warps") == 0) { warps = true; split = false; }
                if (std::strcmp(force, "split") == 0 && split) warps = false;
            }
            if (warps) {
                pl->warps = true;
                pl->block = dim3(32 * N);
                pl->grid = (unsigned)((R + 31) / 32);
            } else if (split) {
                pl->split = true;
                pl->grid = (unsigned)((R * N + mb::SINGLE_THREADS - 1) / mb::SINGLE_THREADS);
            }
        }
    } else if (N > 128 || (std::getenv("MAGPY_B200_CLUSTER_KERNEL") && std::strcmp(std::getenv("MAGPY_B200_CLUSTER_KERNEL"), "big") == 0)) {
        // cluster_big.cu: any cluster size, moments (and the implicit scheme's midpoint iterates) in global memory;
        // MAGPY_B200_CLUSTER_KERNEL=big forces it for smaller clusters (tests: the oracle's dense path is slow beyond 64)
        pl->big = true;
        pl->np = 1;
        pl->block = dim3(mb::CL_LANES, 16);
        pl->grid = (unsigned)((R + mb::CL_LANES - 1) / mb::CL_LANES);
        pl->smem = 0;
    } else if (choose_mma(a, pl)) {
        // cluster_mma.cu: dipolar field as a matrix product on DMMA; geometry set by choose_mma
    } else if (choose_imid_mma(a, pl)) {
        // cluster_mma_imid.cu: the same product inside every quasi-Newton iteration
    } else {
        const uint32_t max_slots = pl->implicit ? 8 : 16;
        // >= 2 own particles per thread reuse every shared-memory moment read of the dipolar sum; the implicit kernel's
        // per-particle Newton work dominates below 8 particles, where one particle per thread (no padding slot) is faster
        int np = (pl->implicit && N < 8) ? 1 : 2;
        while ((N + np - 1) / np > max_slots) np *= 2;
        pl->np = np;
        const uint32_t slots = (N + np - 1) / np;
        pl->block = dim3(mb::CL_LANES, slots);
        pl->grid = (unsigned)((R + mb::CL_LANES - 1) / mb::CL_LANES);
        const size_t moments = (size_t)N * 3 * mb::CL_LANES * sizeof(double);
        const size_t red = (size_t)3 * slots * mb::CL_LANES * sizeof(double);
        const size_t table = (size_t)N * N * 4 * sizeof(double);
        const size_t cap = 227 * 1024;
        pl->layout = 0;
        pl->smem = 2 * moments + red;
        if (pl->implicit) pl->smem += table;   // N <= 64: the implicit kernel always stages the table
        if (!pl->implicit) {   // Heun: stage the pair table in shared memory when it fits (cluster.cu)
            if (2 * moments + red + table <= cap) { pl->layout = 1; pl->smem = 2 * moments + red + table; }
            else if (np == 4 && moments + red + table <= cap) { pl->layout = 2; pl->smem = moments + red + table; }
            else if (np == 4) { np = pl->np = 8; }   // only (8, global table) is instantiated beyond that
            if (pl->layout == 0) {
                const uint32_t sl = (N + pl->np - 1) / pl->np;
                pl->block = dim3(mb::CL_LANES, sl);
                pl->smem = 2 * moments + (size_t)3 * sl * mb::CL_LANES * sizeof(double);
            }
        }
        if (pl->smem > cap) return fail(MAGPY_B200_ERR_BAD_ARG, "cluster too large for shared memory");
    }
    pl->use_table = a->field_shape != MAGPY_B200_FIELD_CONSTANT;
    pl->axis_z = N == 1 && a->axis_stride == 0 && a->anisotropy_axis[0] == 0.0 && a->anisotropy_axis[1] == 0.0 &&
                 a->anisotropy_axis[2] == 1.0;
    choose_k1_variant(pl, a->renorm != 0, a->member_anisotropy || a->member_damping || a->member_field_amplitude);
    // chunking: bound the field table / injected-noise window and the partial-sum buffer
    uint64_t max_steps = 4ull << 20;
    if (const char* env = std::getenv("MAGPY_B200_MAX_CHUNK_STEPS")) {   // test hook: force many small launches
        const long long v = std::atoll(env);
        if (v > 0) max_steps = (uint64_t)v;
    }
    const uint64_t max_partial_doubles = (512ull << 20) / 8;
    uint32_t max_samples_chunk = (uint32_t)std::max<uint64_t>(1, max_partial_doubles / (4ull * pl->grid));
    pl->chunks.clear();
    {
        uint64_t j = 0;
        uint32_t k = 0;
        const uint32_t S = (uint32_t)pl->S;
        while (k < S || j < pl->total_steps) {
            Chunk c{j, j, k, k};
            while (c.k1 < S && c.k1 - c.k0 < max_samples_chunk && pl->target[c.k1] - c.j0 <= max_steps) {
                c.j1 = pl->target[c.k1];
                ++c.k1;
            }
            if (c.k1 == c.k0) {  // next sample is further than max_steps away: advance without sampling
                c.j1 = std::min(pl->target[c.k0], c.j0 + max_steps);
            }
            pl->chunks.push_back(c);
            j = c.j1;
            k = c.k1;
            if (k >= S && j >= pl->total_steps) break;
        }
    }
    for (const Chunk& c : pl->chunks) {
        pl->max_chunk_steps = std::max(pl->max_chunk_steps, c.j1 - c.j0);
        pl->max_chunk_samples = std::max(pl->max_chunk_samples, c.k1 - c.k0);
    }

    // K1b (heun_single_balanced.cu): a single-particle Heun shard of more than one wave of resident CTAs pays for a
    // whole number of warps per SM sub-partition (section 7 of DESIGN.md); cut into (time segment, member block) tasks
    // pulled by a grid of resident CTAs, every SM stays busy to the end.  MAGPY_B200_K1_BALANCE=0|1 overrides.
    pl->bal_segments.assign(pl->chunks.size(), 1);
    if (N == 1 && !pl->implicit && plan_noise(pl) == mb::NOISE_PHILOX_PACKED && pl->k1_min_blocks != mb::K1_LATENCY && !pl->k1_split &&
        !(a->member_anisotropy || a->member_damping || a->member_field_amplitude)) {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pl->device);
        const int per_sm = mb::heun_single_balanced_resident_ctas(pl->use_table, pl->axis_z, a->renorm != 0);
        // four CTAs per SM saturate the issue port (oldest-first warp scheduling); more only hold blocks back — unless every
        // CTA works through many blocks (>= 5: shards of 500,000 members and more), where the tail no longer matters and a
        // fifth warp per sub-partition fills the fixed-latency stalls of the 37-instruction step
        // (profiles/r02_probe_k1_bal_ctas_37op.log, 1e5 steps: 1M members 388.4 ms with 4, 384.9 with 5 or 6; 500k 194.3 /
        // 192.9 / 193.9; 250k 97.2 / 97.7 / 98.6; 125k 49.5 / 50.5 / 51.6)
        int use_per_sm = std::max(1, std::min(per_sm, pl->grid >= (unsigned)(25 * sms) ? 5 : 4));
        if (const char* env = std::getenv("MAGPY_B200_K1_BAL_CTAS")) use_per_sm = std::max(1, std::min(per_sm, std::atoi(env)));
        const bool on_default = pl->grid > (unsigned)(sms * use_per_sm);
        pl->bal_phys = (unsigned)(sms * use_per_sm);
        bool on = on_default;
        if (const char* env = std::getenv("MAGPY_B200_K1_BALANCE")) on = std::atoi(env) != 0;
        if (on) {
            for (size_t ci = 0; ci < pl->chunks.size(); ++ci) {
                const uint64_t span = pl->chunks[ci].j1 - pl->chunks[ci].j0;
                pl->bal_segments[ci] = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(kBalMaxSegments, span / kBalMinSteps));
                if (pl->bal_segments[ci] > 1) pl->k1_balanced = true;
            }
            if (pl->k1_balanced) pl->k1_min_blocks = 1;   // short launches of the same plan use the free-allocation kernel
        }
    }

    // memory budget
    size_t free_b = 0, total_b = 0;
    CU_TRY(cudaMemGetInfo(&free_b, &total_b));
    double need = 8.0 * (3.0 * n * R + (double)n * R /*axis*/ + 4.0 * pl->S + 4.0 * pl->max_chunk_samples * pl->grid +
                         2.0 * pl->max_chunk_steps + (double)N * N * 4);
    if (pl->big) need += 8.0 * (pl->implicit ? 2.0 : 1.0) * (double)n * R;
    if (pl->want_traj) need += 8.0 * 2.0 * (double)pl->S * n * R;
    if (pl->injected) need += 8.0 * 2.0 * (double)pl->total_steps * n * R;
    if (need > 0.9 * (double)free_b) {   // blocks cached by the pool count as used: give them back and look again
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, pl->device) == cudaSuccess) {
            cudaDeviceSynchronize();
            cudaMemPoolTrimTo(pool, 0);
        }
        CU_TRY(cudaMemGetInfo(&free_b, &total_b));
    }
    if (need > 0.9 * (double)free_b)
        return fail(MAGPY_B200_ERR_NOMEM, "request needs %.1f GB of device memory, %.1f GB free", need / 1e9, free_b / 1e9);

    rc = enable_pool(pl->device);
    if (rc) return rc;
    CU_TRY(cudaStreamCreateWithFlags(&pl->stream, cudaStreamNonBlocking));
    CU_TRY(cudaEventCreate(&pl->ev_begin));
    CU_TRY(cudaEventCreate(&pl->ev_end));
    pl->ev_k.resize(2 * pl->chunks.size());
    for (auto& e : pl->ev_k) CU_TRY(cudaEventCreate(&e));

    // uploads
    CU_TRY(pl->d_state0.alloc(n * R, pl->stream));
    CU_TRY(pl->d_state.alloc(n * R, pl->stream));
    if (pl->big) CU_TRY(pl->d_state_t.alloc(n * R, pl->stream));
    if (pl->big && pl->implicit) CU_TRY(pl->d_state_u.alloc(n * R, pl->stream));
    CU_TRY(pl->d_kred.alloc(N, pl->stream));
    pl->mp = a->member_anisotropy || a->member_damping || a->member_field_amplitude;
    const bool member_radii = !pl->mp && (a->radius_stride != 0 || a->member_temperature != nullptr);   // N = 1: sigma per member
    CU_TRY(pl->d_sig.alloc((member_radii || pl->mp) ? R : N, pl->stream));
    CU_TRY(pl->d_seeds.alloc(R, pl->stream));
    CU_TRY(pl->d_target.alloc(pl->S, pl->stream));
    CU_TRY(pl->d_sums.alloc(pl->S * 4, pl->stream));
    CU_TRY(pl->d_partial.alloc((size_t)4 * pl->max_chunk_samples * pl->grid, pl->stream));
    CU_TRY(pl->d_newton.alloc(3, pl->stream));
    if (pl->use_table) CU_TRY(pl->d_tab.alloc(2 * (std::max<uint64_t>(1, pl->max_chunk_steps) + 2 * kMpMargin), pl->stream));
    CU_TRY(cudaMemcpyAsync(pl->d_kred.p, rd.k_red.data(), N * 8, cudaMemcpyHostToDevice, pl->stream));
    std::vector<double> member_sigma;
    if (member_radii) {
        // lib/simulation.cpp:531-535 per member: sigma = sqrt(alpha KB T / (K_av V_r) / (1 + alpha^2)), same order as reduce_units
        member_sigma.resize(R);
        for (uint64_t r = 0; r < R; ++r) {
            const double rad = a->radius[a->radius_stride ? r : 0];
            const double vol = 4.0 / 3.0 * M_PI * rad * rad * rad;
            const double T = a->member_temperature ? a->member_temperature[r] : a->temperature;
            member_sigma[r] = std::sqrt(a->damping * kKB * T / (rd.K_av * vol) / (1 + a->damping * a->damping));
        }
        CU_TRY(cudaMemcpyAsync(pl->d_sig.p, member_sigma.data(), R * 8, cudaMemcpyHostToDevice, pl->stream));
        CU_TRY(cudaStreamSynchronize(pl->stream));
        pl->h2d += R * 8;
    } else if (!pl->mp) {
        CU_TRY(cudaMemcpyAsync(pl->d_sig.p, rd.sigma.data(), N * 8, cudaMemcpyHostToDevice, pl->stream));
    }
    std::vector<double> mp_host;
    if (pl->mp) {
        // lib/simulation.cpp:498-549 member by member (N = 1: V_av = V, K_av = K, k_red = v_red = 1), same evaluation
        // order as reduce_units: rows alpha | dt_red | h0 | sampling interval, and sigma into d_sig
        mp_host.resize(4 * R);
        member_sigma.resize(R);
        for (uint64_t r = 0; r < R; ++r) {
            const double rad = a->radius[a->radius_stride ? r : 0];
            const double K = a->member_anisotropy ? a->member_anisotropy[r] : a->anisotropy[0];
            const double al = a->member_damping ? a->member_damping[r] : a->damping;
            const double T = a->member_temperature ? a->member_temperature[r] : a->temperature;
            const double H0 = a->member_field_amplitude ? a->member_field_amplitude[r] : a->field_amplitude;
            const double vol = 4.0 / 3.0 * M_PI * rad * rad * rad;
            const double K_av = K / 1;
            const double H_k = 2 * K_av / kMU0 / a->magnetisation;
            const double tau = kGYROMAG * kMU0 * H_k / (1 + al * al);
            const double dtr = a->time_step * tau, Tr = a->end_time * tau;
            if (!(dtr > 0.0) || !std::isfinite(dtr) || !std::isfinite(Tr))
                return fail(MAGPY_B200_ERR_BAD_ARG, "member %llu: non-finite reduced time step", (unsigned long long)r);
            mp_host[r] = al;
            mp_host[R + r] = dtr;
            mp_host[2 * R + r] = H0 / H_k;
            mp_host[3 * R + r] = Tr / (pl->S - 1);
            member_sigma[r] = std::sqrt(al * kKB * T / (K_av * vol) / (1 + al * al));
        }
        CU_TRY(pl->d_mp.alloc(4 * R, pl->stream));
        CU_TRY(pl->d_member_j.alloc(R, pl->stream));
        CU_TRY(cudaMemcpyAsync(pl->d_mp.p, mp_host.data(), 4 * R * 8, cudaMemcpyHostToDevice, pl->stream));
        CU_TRY(cudaMemcpyAsync(pl->d_sig.p, member_sigma.data(), R * 8, cudaMemcpyHostToDevice, pl->stream));
        CU_TRY(cudaStreamSynchronize(pl->stream));
        pl->h2d += 5 * R * 8;
    }
    CU_TRY(cudaMemcpyAsync(pl->d_target.p, pl->target.data(), pl->S * 8, cudaMemcpyHostToDevice, pl->stream));
    pl->h2d += N * 16 + pl->S * 8;
    std::vector<uint32_t> seg_k;
    if (pl->k1_balanced) {
        // first sample of the chunk whose state index is >= the first step of each time segment
        seg_k.assign(pl->chunks.size() * kBalMaxSegments, 0);
        for (size_t ci = 0; ci < pl->chunks.size(); ++ci) {
            const Chunk& c = pl->chunks[ci];
            const uint64_t span = c.j1 - c.j0, ns = pl->bal_segments[ci];
            uint32_t k = c.k0;
            for (uint64_t sgm = 0; sgm < ns; ++sgm) {
                const uint64_t ja = c.j0 + span * sgm / ns;
                while (k < c.k1 && pl->target[k] < ja) ++k;
                seg_k[ci * kBalMaxSegments + sgm] = k;
            }
        }
        CU_TRY(pl->d_bal_seg_k.alloc(seg_k.size(), pl->stream));
        CU_TRY(pl->d_bal_sync.alloc(1 + (size_t)pl->grid, pl->stream));
        CU_TRY(cudaMemcpyAsync(pl->d_bal_seg_k.p, seg_k.data(), seg_k.size() * 4, cudaMemcpyHostToDevice, pl->stream));
        CU_TRY(cudaStreamSynchronize(pl->stream));
        pl->h2d += seg_k.size() * 4;
    }
    std::vector<int64_t> zero_seeds;
    const int64_t* seeds = a->seeds;
    if (!seeds) {
        zero_seeds.assign(R, 0);
        seeds = zero_seeds.data();
    }
    CU_TRY(cudaMemcpyAsync(pl->d_seeds.p, seeds, R * 8, cudaMemcpyHostToDevice, pl->stream));
    pl->h2d += R * 8;
    std::vector<uint32_t> member_idx;
    if (a->member_index) {
        member_idx.resize(R);
        for (uint64_t r = 0; r < R; ++r) member_idx[r] = (uint32_t)a->member_index[r];
        CU_TRY(pl->d_member_idx.alloc(R, pl->stream));
        CU_TRY(cudaMemcpyAsync(pl->d_member_idx.p, member_idx.data(), R * 4, cudaMemcpyHostToDevice, pl->stream));
        CU_TRY(cudaStreamSynchronize(pl->stream));
        pl->h2d += R * 4;
    }
    pl->comm = a->comm;
    pl->unit_field = a->member_field_amplitude != nullptr;

    // per-member arrays arrive [R][n]; the device wants [n][R]
    {
        size_t stage = n * R;
        if (pl->injected) stage = std::max<size_t>(stage, (size_t)pl->total_steps * n * R);
        if (pl->want_traj) stage = std::max<size_t>(stage, (size_t)pl->S * n * R);
        CU_TRY(pl->d_stage.alloc(stage, pl->stream));
    }
    if (a->m0_stride) {
        CU_TRY(cudaMemcpyAsync(pl->d_stage.p, a->magnetisation_direction, n * R * 8, cudaMemcpyHostToDevice, pl->stream));
        pl->h2d += n * R * 8;
        rc = launch_transpose(pl, pl->d_stage.p, pl->d_state0.p, 1, R, n, 0, n, 0, R, 1.0);
        if (rc) return rc;
    } else {
        // shared initial state: upload the n values once and replicate them on the device
        CU_TRY(cudaMemcpyAsync(pl->d_stage.p, a->magnetisation_direction, n * 8, cudaMemcpyHostToDevice, pl->stream));
        pl->h2d += n * 8;
        LAUNCH_TRY(mb::launch_broadcast_rows(pl->d_stage.p, pl->d_state0.p, n, R, pl->stream));
    }
    if (a->axis_stride) {
        CU_TRY(pl->d_axis.alloc(n * R, pl->stream));
        CU_TRY(cudaMemcpyAsync(pl->d_stage.p, a->anisotropy_axis, n * R * 8, cudaMemcpyHostToDevice, pl->stream));
        pl->h2d += n * R * 8;
        rc = launch_transpose(pl, pl->d_stage.p, pl->d_axis.p, 1, R, n, 0, n, 0, R, 1.0);
        if (rc) return rc;
    } else {
        CU_TRY(pl->d_axis.alloc(n, pl->stream));
        CU_TRY(cudaMemcpyAsync(pl->d_axis.p, a->anisotropy_axis, n * 8, cudaMemcpyHostToDevice, pl->stream));
        pl->h2d += n * 8;
    }
    if (pl->injected) {
        // host [R][steps][n] -> device [steps][n][R]
        const uint64_t T = pl->total_steps;
        CU_TRY(pl->d_dW.alloc(std::max<uint64_t>(1, T * n * R), pl->stream));
        if (T > 0) {
            if (a->injected_steps == T) {
                CU_TRY(cudaMemcpyAsync(pl->d_stage.p, a->injected_dw, T * n * R * 8, cudaMemcpyHostToDevice, pl->stream));
            } else {
                CU_TRY(cudaMemcpy2DAsync(pl->d_stage.p, T * n * 8, a->injected_dw, a->injected_steps * n * 8, T * n * 8, R,
                                         cudaMemcpyHostToDevice, pl->stream));
            }
            pl->h2d += T * n * R * 8;
            rc = launch_transpose(pl, pl->d_stage.p, pl->d_dW.p, 1, R, T * n, 0, T * n, 0, R, 1.0);
            if (rc) return rc;
        }
    }
    // dipolar pair table (lib/distances.cpp:20-113, lib/simulation.cpp:577-583, lib/field.cpp:187-225)
    if (N > 1) {
        std::vector<double> tab((size_t)N * N * 4, 0.0);
        const double lscale = std::pow(rd.V_av, 1. / 3);
        for (uint32_t i = 0; i < N; ++i)
            for (uint32_t jx = 0; jx < N; ++jx) {
                if (i == jx) continue;
                double d[3];
                for (int c = 0; c < 3; ++c) d[c] = a->location[3 * jx + c] - a->location[3 * i + c];
                const double mag = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                const double cube = std::pow(mag / lscale, 3);
                double* t = &tab[((size_t)i * N + jx) * 4];
                // sqrt(3) r_hat: (m.t) t = 3 (m.r_hat) r_hat (lib/field.cpp:221-224)
                const double s3 = 1.7320508075688772;
                t[0] = s3 * (d[0] / mag); t[1] = s3 * (d[1] / mag); t[2] = s3 * (d[2] / mag);
                t[3] = rd.dip_pre * (rd.v_red[jx] / cube);
            }
        CU_TRY(pl->d_dip.alloc(tab.size(), pl->stream));
        CU_TRY(cudaMemcpyAsync(pl->d_dip.p, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice, pl->stream));
        CU_TRY(cudaStreamSynchronize(pl->stream));
        pl->h2d += tab.size() * 8;
    }
    if (pl->mma || pl->imid_mma) {
        // symmetric dipolar matrix D[(i,a),(j,b)] = c_dip / cube_ij (3 r_a r_b - delta_ab) in the row order
        // 24 (p / 8) + 8 a + p % 8, upper 24 x 24 blocks only, swizzled inside a block (cluster_mma.cu)
        const uint32_t G = pl->G;
        std::vector<double> dm((size_t)G * (G + 1) / 2 * 576, 0.0);
        const double lscale = std::pow(rd.V_av, 1. / 3);
        for (uint32_t i = 0; i < N; ++i)
            for (uint32_t jx = 0; jx < N; ++jx) {
                if (i == jx || jx / 8 < i / 8) continue;
                double d[3];
                for (int c = 0; c < 3; ++c) d[c] = a->location[3 * jx + c] - a->location[3 * i + c];
                const double mag = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                const double cube = std::pow(mag / lscale, 3);
                const double pre = rd.dip_pre / cube;
                const uint32_t pg = i / 8, kg = jx / 8;
                double* blk = &dm[((size_t)pg * G - (size_t)pg * (pg - 1) / 2 + (kg - pg)) * 576];
                for (int aa = 0; aa < 3; ++aa)
                    for (int bb = 0; bb < 3; ++bb) {
                        const uint32_t row = 8 * aa + i % 8, col = 8 * bb + jx % 8;
                        blk[row * 24 + (col ^ (((row >> 1) & 1) << 2))] =
                            pre * (3.0 * (d[aa] / mag) * (d[bb] / mag) - (aa == bb ? 1.0 : 0.0));
                    }
            }
        CU_TRY(pl->d_dmat.alloc(dm.size(), pl->stream));
        CU_TRY(pl->d_vred.alloc(N, pl->stream));
        CU_TRY(cudaMemcpyAsync(pl->d_dmat.p, dm.data(), dm.size() * 8, cudaMemcpyHostToDevice, pl->stream));
        CU_TRY(cudaMemcpyAsync(pl->d_vred.p, rd.v_red.data(), N * 8, cudaMemcpyHostToDevice, pl->stream));
        CU_TRY(cudaStreamSynchronize(pl->stream));
        pl->h2d += dm.size() * 8 + N * 8;
    }
    if (pl->want_traj) CU_TRY(pl->d_traj.alloc((size_t)pl->S * n * R, pl->stream));
    CU_TRY(cudaStreamSynchronize(pl->stream));

    mb::RunParams& P = pl->base;
    P.R = R;
    P.N = N;
    P.renorm = a->renorm;
    P.interactions = (a->interactions != 0 && N > 1) ? 1 : 0;
    P.alpha = a->damping;
    P.dt = rd.dt;
    P.sqrt_dt = std::sqrt(rd.dt);
    P.eps = a->implicit_tol;
    P.quirk_zero = (N >= 2 && a->axis_stride == 0 && a->anisotropy_axis[0] == 0.0 && a->anisotropy_axis[1] == 0.0) ? 1 : 0;
    P.newton_exact = (a->use_implicit && a->implicit_newton == MAGPY_B200_NEWTON_EXACT) ? 1 : 0;
    P.clampA = std::sqrt(2 * 1000.0 * std::abs(std::log(rd.dt)));  // lib/integrators.cpp:598-599
    P.h_const = rd.h0;
    P.k_red = pl->d_kred.p;
    P.half_kdt0 = 0.5 * (rd.k_red[0] * rd.dt);
    P.half_dt = 0.5 * rd.dt;
    P.sig = pl->d_sig.p;
    P.sig_rs = (pl->mp || a->radius_stride != 0 || a->member_temperature != nullptr) ? 1 : 0;
    if (pl->mp) {
        P.mp_alpha = pl->d_mp.p;
        P.mp_dt = pl->d_mp.p + R;
        P.mp_h0 = pl->d_mp.p + 2 * R;
        P.mp_Ts = pl->d_mp.p + 3 * R;
        P.member_j = pl->d_member_j.p;
    }
    P.dip = pl->d_dip.p;
    P.dmat = pl->d_dmat.p;
    P.v_red = pl->d_vred.p;
    P.G = pl->G;
    P.mma_full = pl->mma_full;
    P.mma_dglobal = pl->mma_dglobal ? 1u : 0u;
    P.mma_mono = std::all_of(rd.v_red.begin(), rd.v_red.end(), [](double v) { return v == 1.0; }) ? 1u : 0u;
    P.mma_tail = pl->mma_tail;
    P.axis = pl->d_axis.p;
    P.axis_cs = a->axis_stride ? R : 1;
    P.axis_rs = a->axis_stride ? 1 : 0;
    P.seeds = pl->d_seeds.p;
    P.stream_offset = a->stream_offset;
    P.member_idx = pl->d_member_idx.p;
    P.coarsen_log2 = a->noise_coarsen_log2;
    P.philox_m0 = 0xD2511F53u;
    P.philox_m1 = 0xCD9E8D57u;
    P.bm_mask_r = 0x007fffffu;
    P.bm_mask_a = 0x007fffe0u;
    P.state = pl->d_state.p;
    P.state_t = pl->d_state_t.p;
    P.state_u = pl->d_state_u.p;
    P.target = pl->d_target.p;
    P.field_tab = pl->d_tab.p;
    P.dW = pl->d_dW.p;
    P.dW_j0 = 0;
    P.traj = pl->d_traj.p;
    P.partial = pl->d_partial.p;
    P.newton = pl->d_newton.p;
    return MAGPY_B200_OK;
}

int plan_run(magpy_b200_plan* pl) {
    CU_TRY(cudaSetDevice(pl->device));
    const uint64_t nR = (uint64_t)pl->n * pl->R;
    CU_TRY(cudaEventRecord(pl->ev_begin, pl->stream));
    CU_TRY(cudaMemcpyAsync(pl->d_state.p, pl->d_state0.p, nR * 8, cudaMemcpyDeviceToDevice, pl->stream));
    CU_TRY(cudaMemsetAsync(pl->d_newton.p, 0, 3 * sizeof(unsigned long long), pl->stream));
    CU_TRY(cudaMemsetAsync(pl->d_sums.p, 0, pl->S * 4 * 8, pl->stream));
    if (pl->mp) CU_TRY(cudaMemsetAsync(pl->d_member_j.p, 0, pl->R * 4, pl->stream));
    const double second = pl->implicit ? pl->red.dt / 2 : pl->red.dt;
    size_t ci = 0;
    for (const Chunk& c : pl->chunks) {
        mb::RunParams P = pl->base;
        P.j0 = c.j0; P.j1 = c.j1; P.k0 = c.k0; P.k1 = c.k1;
        const uint64_t ns = c.j1 - c.j0;
        if (pl->mp) {   // unit waveform, a margin of steps on either side (per-member schedules)
            P.tab_j0 = c.j0 > kMpMargin ? c.j0 - kMpMargin : 0;
            if (pl->use_table)
                LAUNCH_TRY(mb::launch_field_table(pl->d_tab.p, P.tab_j0, c.j1 + kMpMargin - P.tab_j0, pl->red.dt, second,
                                                  pl->field_shape, 1.0, pl->red.f, pl->stream));
        } else if (pl->use_table && ns > 0) {
            LAUNCH_TRY(mb::launch_field_table(pl->d_tab.p, c.j0, ns, pl->red.dt, second, pl->field_shape, pl->red.h0,
                                              pl->red.f, pl->stream));
        }
        const bool balanced = pl->k1_balanced && pl->bal_segments[ci] > 1;
        if (balanced) {   // task counter and per-block flags back to zero, this chunk's first-sample table
            CU_TRY(cudaMemsetAsync(pl->d_bal_sync.p, 0, (size_t)(1 + pl->grid) * sizeof(unsigned int), pl->stream));
            P.bal_vctas = pl->grid;
            P.bal_segments = pl->bal_segments[ci];
            P.bal_counter = pl->d_bal_sync.p;
            P.bal_flags = pl->d_bal_sync.p + 1;
            P.bal_seg_k = pl->d_bal_seg_k.p + ci * kBalMaxSegments;
        }
        CU_TRY(cudaEventRecord(pl->ev_k[2 * ci], pl->stream));
        if (balanced) {
            const uint64_t tasks = (uint64_t)pl->grid * P.bal_segments;
            LAUNCH_TRY(mb::launch_heun_single_balanced(pl->use_table, pl->axis_z, (unsigned)std::min<uint64_t>(tasks, pl->bal_phys),
                                                       pl->stream, P));
        } else {
            int rc = launch_integrate(pl, P);
            if (rc) return rc;
        }
        CU_TRY(cudaEventRecord(pl->ev_k[2 * ci + 1], pl->stream));
        if (c.k1 > c.k0) {
            LAUNCH_TRY(mb::launch_reduce_partials(pl->d_partial.p, pl->d_sums.p, c.k0, c.k1 - c.k0, pl->grid, pl->stream));
        }
        ++ci;
    }
    if (pl->comm) {   // the one collective of the multi-GPU path: sums over all ranks' members, in place, on this stream
        int rc = mbh::comm_allreduce_device(pl->comm, pl->d_sums.p, pl->S * 4, MAGPY_B200_COMM_SUM, pl->stream);
        if (rc) return rc;
    }
    CU_TRY(cudaEventRecord(pl->ev_end, pl->stream));
    pl->ran = true;
    return MAGPY_B200_OK;
}

int plan_sync(magpy_b200_plan* pl, magpy_b200_stats* st) {
    CU_TRY(cudaSetDevice(pl->device));
    CU_TRY(cudaStreamSynchronize(pl->stream));
    if (!st) return MAGPY_B200_OK;
    std::memset(st, 0, sizeof *st);
    st->steps_per_member = pl->total_steps;
    st->particle_steps = pl->total_steps * pl->R * pl->N;
    st->kernel_launches = pl->launches;
    st->h2d_bytes = pl->h2d;
    st->d2h_bytes = pl->d2h;
    st->kernel_family = pl->N == 1 ? (pl->implicit ? MAGPY_B200_KERNEL_IMID_SINGLE
                                      : MAGPY_B200_KERNEL_HEUN_SINGLE)
                        : pl->warps ? MAGPY_B200_KERNEL_IMID_WARPS
                        : pl->split ? MAGPY_B200_KERNEL_IMID_SPLIT
                        : pl->small ? (pl->implicit ? MAGPY_B200_KERNEL_IMID_SMALL : MAGPY_B200_KERNEL_HEUN_SMALL)
                        : pl->big   ? (pl->implicit ? MAGPY_B200_KERNEL_IMID_CLUSTER_BIG : MAGPY_B200_KERNEL_HEUN_CLUSTER_BIG)
                        : pl->mma   ? MAGPY_B200_KERNEL_HEUN_CLUSTER_MMA
                        : pl->imid_mma ? MAGPY_B200_KERNEL_IMID_CLUSTER_MMA
                                    : (pl->implicit ? MAGPY_B200_KERNEL_IMID_CLUSTER : MAGPY_B200_KERNEL_HEUN_CLUSTER);
    st->kernel_variant = (pl->N == 1 && !pl->implicit) ? (pl->k1_balanced ? 200u : pl->k1_split ? (uint64_t)mb::K1_SPLIT : (uint64_t)pl->k1_min_blocks) : 0;
    if (pl->ran) {
        float ms = 0.f;
        CU_TRY(cudaEventElapsedTime(&ms, pl->ev_begin, pl->ev_end));
        st->device_ms = ms;
        double acc = 0.0;
        for (size_t i = 0; i < pl->chunks.size(); ++i) {
            CU_TRY(cudaEventElapsedTime(&ms, pl->ev_k[2 * i], pl->ev_k[2 * i + 1]));
            acc += ms;
        }
        st->integrate_ms = acc;
        unsigned long long nw[3];
        CU_TRY(cudaMemcpy(nw, pl->d_newton.p, sizeof nw, cudaMemcpyDeviceToHost));
        st->newton_iterations = nw[0];
        st->newton_max_iterations = nw[1];
        st->newton_failures = nw[2];
    }
    return MAGPY_B200_OK;
}

int plan_fetch(magpy_b200_plan* pl, double* out_time, double* out_field, double* out_traj, double* out_sums,
               double* out_final) {
    CU_TRY(cudaSetDevice(pl->device));
    const Reduced& rd = pl->red;
    const uint64_t S = pl->S, n = pl->n, R = pl->R;
    const double Ts = rd.T / (S - 1);
    // lib/simulation.cpp:398-401 then :613-616
    if (out_time)
        for (uint64_t k = 0; k < S; ++k) out_time[k] = ((double)(unsigned int)k * Ts) / rd.tau;
    if (out_field)
        for (uint64_t k = 0; k < S; ++k)
            out_field[k] = pl->unit_field ? applied_field(pl->field_shape, (double)(unsigned int)k * Ts, 1.0, rd.f)
                                          : applied_field(pl->field_shape, (double)(unsigned int)k * Ts, rd.h0, rd.f) * rd.H_k;
    if (out_sums) {
        CU_TRY(cudaMemcpyAsync(out_sums, pl->d_sums.p, S * 4 * 8, cudaMemcpyDeviceToHost, pl->stream));
        CU_TRY(cudaStreamSynchronize(pl->stream));
        pl->d2h += S * 4 * 8;
        for (uint64_t k = 0; k < S; ++k) {
            out_sums[4 * k + 0] *= pl->Ms;
            out_sums[4 * k + 1] *= pl->Ms;
            out_sums[4 * k + 2] *= pl->Ms;
            out_sums[4 * k + 3] *= pl->Ms * pl->Ms;
        }
    }
    if (out_final) {
        int rc = launch_transpose(pl, pl->d_state.p, pl->d_stage.p, 1, n, R, 0, R, 0, n, pl->Ms);
        if (rc) return rc;
        CU_TRY(cudaMemcpyAsync(out_final, pl->d_stage.p, n * R * 8, cudaMemcpyDeviceToHost, pl->stream));
        CU_TRY(cudaStreamSynchronize(pl->stream));
        pl->d2h += n * R * 8;
    }
    if (out_traj) {
        if (!pl->want_traj) return fail(MAGPY_B200_ERR_BAD_ARG, "plan was created without out_trajectories");
        // device [S][n][R] -> host [R][n][S], scaled to A/m (lib/simulation.cpp:617-620)
        int rc = MAGPY_B200_OK;
        if ((uint64_t)n * S <= 65535ull * 128) {   // dedicated kernel: full 256 B lines on both sides (service.cu)
            LAUNCH_TRY(mb::launch_traj_fetch(pl->d_traj.p, pl->d_stage.p, R, (uint32_t)n, (uint32_t)S, pl->Ms, pl->stream));
        } else {
            rc = launch_transpose(pl, pl->d_traj.p, pl->d_stage.p, n, S, R, R, n * R, S, n * S, pl->Ms);
        }
        if (rc) return rc;
        CU_TRY(cudaMemcpyAsync(out_traj, pl->d_stage.p, S * n * R * 8, cudaMemcpyDeviceToHost, pl->stream));
        CU_TRY(cudaStreamSynchronize(pl->stream));
        pl->d2h += S * n * R * 8;
    }
    return MAGPY_B200_OK;
}

}  // namespace

// =======================================================================================
// C ABI
// =======================================================================================
extern "C" {

int magpy_b200_abi_version(void) { return MAGPY_B200_ABI_VERSION; }
const char* magpy_b200_last_error(void) { return mbh::g_error.c_str(); }

int magpy_b200_device_count(int* count) {
    if (!count) return fail(MAGPY_B200_ERR_BAD_ARG, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    *count = n;
    return MAGPY_B200_OK;
}

double magpy_b200_get_KB(void) { return kKB; }
double magpy_b200_get_mu0(void) { return kMU0; }
double magpy_b200_get_gamma(void) { return kGYROMAG; }

int magpy_b200_reduce_units(const double* radius, const double* anisotropy, size_t N, double Ms, double alpha,
                            double T, double dt, double t_end, double H0, double f, double* k_red, double* v_red,
                            double* sigma, double* out) {
    if (!radius || !anisotropy || N == 0 || !out) return fail(MAGPY_B200_ERR_BAD_ARG, "bad arguments");
    Reduced r;
    reduce_units(radius, anisotropy, N, Ms, alpha, T, dt, t_end, H0, f, r);
    for (size_t i = 0; i < N; ++i) {
        if (k_red) k_red[i] = r.k_red[i];
        if (v_red) v_red[i] = r.v_red[i];
        if (sigma) sigma[i] = r.sigma[i];
    }
    out[0] = r.V_av; out[1] = r.K_av; out[2] = r.H_k; out[3] = r.tau; out[4] = r.dt;
    out[5] = r.T; out[6] = r.h0; out[7] = r.f; out[8] = r.dip_pre;
    return MAGPY_B200_OK;
}

int magpy_b200_schedule(double dt_red, double t_end_red, size_t S, uint64_t* cum) {
    if (S < 2 || !cum || !(dt_red > 0)) return fail(MAGPY_B200_ERR_BAD_ARG, "bad arguments");
    std::vector<uint64_t> c;
    if (build_schedule(dt_red, t_end_red, S, c)) return fail(MAGPY_B200_ERR_BAD_ARG, "step count exceeds 32 bits");
    std::copy(c.begin(), c.end(), cum);
    return MAGPY_B200_OK;
}

int magpy_b200_release_cached_memory(int device) {
    int rc = select_device(device);
    if (rc) return rc;
    cudaMemPool_t pool;
    CU_TRY(cudaDeviceGetDefaultMemPool(&pool, device));
    CU_TRY(cudaDeviceSynchronize());
    CU_TRY(cudaMemPoolTrimTo(pool, 0));
    return MAGPY_B200_OK;
}

// ---- pinned host buffers for output arrays (cached by size) ----------------------------------------------------
namespace {
struct PinnedCache {
    std::mutex mu;
    std::multimap<size_t, void*> free_blocks;     // rounded size -> block
    std::unordered_map<void*, size_t> live;       // blocks handed out
    size_t cached_bytes = 0;
    size_t limit() const {
        if (const char* env = std::getenv("MAGPY_B200_PINNED_CACHE_MB")) return (size_t)std::max(0ll, std::atoll(env)) << 20;
        return (size_t)8192 << 20;
    }
} g_pinned;
size_t pinned_round(size_t bytes) {
    // 64 KiB granules up to 16 MiB, then 1/8 of the leading power of two: a freed block serves every request within 12.5 %
    size_t g = 64 << 10;
    while ((g << 3) < bytes) g <<= 1;
    return (bytes + g - 1) / g * g;
}
}  // namespace

int magpy_b200_host_alloc(size_t bytes, void** ptr) {
    if (!ptr || bytes == 0) return fail(MAGPY_B200_ERR_BAD_ARG, "bad arguments");
    *ptr = nullptr;
    const size_t want = pinned_round(bytes);
    {
        std::lock_guard<std::mutex> lock(g_pinned.mu);
        auto it = g_pinned.free_blocks.find(want);
        if (it != g_pinned.free_blocks.end()) {
            *ptr = it->second;
            g_pinned.cached_bytes -= want;
            g_pinned.free_blocks.erase(it);
            g_pinned.live[*ptr] = want;
            return MAGPY_B200_OK;
        }
    }
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(MAGPY_B200_ERR_NO_DEVICE, "no CUDA device available: no page-locked memory");
    }
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocPortable);
    if (e != cudaSuccess) {   // give the cache back and try once more
        cudaGetLastError();
        magpy_b200_host_cache_release();
        e = cudaHostAlloc(&p, want, cudaHostAllocPortable);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(MAGPY_B200_ERR_NOMEM, "cudaHostAlloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    std::lock_guard<std::mutex> lock(g_pinned.mu);
    g_pinned.live[p] = want;
    *ptr = p;
    return MAGPY_B200_OK;
}

int magpy_b200_host_free(void* ptr) {
    if (!ptr) return MAGPY_B200_OK;
    size_t size = 0;
    bool keep = false;
    {
        std::lock_guard<std::mutex> lock(g_pinned.mu);
        auto it = g_pinned.live.find(ptr);
        if (it == g_pinned.live.end()) return fail(MAGPY_B200_ERR_BAD_ARG, "pointer was not returned by magpy_b200_host_alloc");
        size = it->second;
        g_pinned.live.erase(it);
        if (g_pinned.cached_bytes + size <= g_pinned.limit()) {
            g_pinned.free_blocks.emplace(size, ptr);
            g_pinned.cached_bytes += size;
            keep = true;
        }
    }
    if (!keep) cudaFreeHost(ptr);
    return MAGPY_B200_OK;
}

int magpy_b200_host_cache_release(void) {
    std::vector<void*> blocks;
    {
        std::lock_guard<std::mutex> lock(g_pinned.mu);
        for (auto& kv : g_pinned.free_blocks) blocks.push_back(kv.second);
        g_pinned.free_blocks.clear();
        g_pinned.cached_bytes = 0;
    }
    for (void* p : blocks) cudaFreeHost(p);
    return MAGPY_B200_OK;
}

int magpy_b200_plan_create(const magpy_b200_ensemble* args, magpy_b200_plan** plan) {
    if (!plan) return fail(MAGPY_B200_ERR_BAD_ARG, "plan is NULL");
    *plan = nullptr;
    magpy_b200_plan* pl = new (std::nothrow) magpy_b200_plan();
    if (!pl) return fail(MAGPY_B200_ERR_NOMEM, "out of host memory");
    int rc;
    try {
        rc = plan_build(args, pl);
    } catch (const std::bad_alloc&) {
        rc = fail(MAGPY_B200_ERR_NOMEM, "out of host memory");
    } catch (const std::exception& e) {
        rc = fail(MAGPY_B200_ERR_BAD_ARG, "%s", e.what());
    }
    if (rc) {
        delete pl;
        return rc;
    }
    *plan = pl;
    return MAGPY_B200_OK;
}

int magpy_b200_plan_run(magpy_b200_plan* plan) {
    if (!plan) return fail(MAGPY_B200_ERR_BAD_ARG, "plan is NULL");
    return plan_run(plan);
}

int magpy_b200_plan_sync(magpy_b200_plan* plan, magpy_b200_stats* stats) {
    if (!plan) return fail(MAGPY_B200_ERR_BAD_ARG, "plan is NULL");
    return plan_sync(plan, stats);
}

int magpy_b200_plan_fetch(magpy_b200_plan* plan, double* out_time, double* out_field, double* out_trajectories,
                          double* out_sums, double* out_final) {
    if (!plan) return fail(MAGPY_B200_ERR_BAD_ARG, "plan is NULL");
    try {
        return plan_fetch(plan, out_time, out_field, out_trajectories, out_sums, out_final);
    } catch (const std::exception& e) {
        return fail(MAGPY_B200_ERR_NOMEM, "%s", e.what());
    }
}

int magpy_b200_plan_sums_device_ptr(magpy_b200_plan* plan, void** dptr, size_t* n_doubles) {
    if (!plan || !dptr) return fail(MAGPY_B200_ERR_BAD_ARG, "bad arguments");
    *dptr = plan->d_sums.p;
    if (n_doubles) *n_doubles = plan->S * 4;
    return MAGPY_B200_OK;
}

int magpy_b200_plan_destroy(magpy_b200_plan* plan) {
    delete plan;
    return MAGPY_B200_OK;
}

int magpy_b200_simulate_ensemble(const magpy_b200_ensemble* args, magpy_b200_stats* stats) {
    using clk = std::chrono::steady_clock;
    auto ms = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    const auto t0 = clk::now();
    magpy_b200_plan* pl = nullptr;
    int rc = magpy_b200_plan_create(args, &pl);
    if (rc) return rc;
    const auto t1 = clk::now();
    rc = plan_run(pl);
    if (!rc) rc = plan_sync(pl, nullptr);
    const auto t2 = clk::now();
    if (!rc)
        rc = magpy_b200_plan_fetch(pl, args->out_time, args->out_field, args->out_trajectories, args->out_sums,
                                   args->out_final);
    const auto t3 = clk::now();
    magpy_b200_stats st;
    if (!rc) rc = plan_sync(pl, &st);
    delete pl;
    if (!rc && stats) {
        st.host_setup_ms = ms(t0, t1);
        st.host_run_ms = ms(t1, t2);
        st.host_fetch_ms = ms(t2, t3);
        st.host_total_ms = ms(t0, clk::now());
        *stats = st;
    }
    return rc;
}

int magpy_b200_simulate_ensemble_multi(const magpy_b200_ensemble* args, const int* devices, int n_devices,
                                       magpy_b200_stats* stats) {
    if (!args) return fail(MAGPY_B200_ERR_BAD_ARG, "args is NULL");
    if (!devices || n_devices < 1) return fail(MAGPY_B200_ERR_BAD_ARG, "devices must list at least one CUDA device");
    if (args->comm && n_devices > 1)
        return fail(MAGPY_B200_ERR_BAD_ARG, "a communicator belongs to one device: use either `comm` (one process per GPU) or a device list");
    if (n_devices == 1) {
        magpy_b200_ensemble one = *args;
        one.device = devices[0];
        return magpy_b200_simulate_ensemble(&one, stats);
    }
    const uint64_t R = args->n_members, n = 3ull * args->n_particles, S = args->max_samples;
    const uint64_t per = (R + n_devices - 1) / n_devices;   // same split as magpy_b200/sharding.py
    std::vector<magpy_b200_plan*> plans;
    std::vector<magpy_b200_ensemble> shard;
    std::vector<std::vector<double>> part;
    int rc = MAGPY_B200_OK;
    try {
        for (int g = 0; g < n_devices && !rc; ++g) {
            const uint64_t lo = std::min<uint64_t>((uint64_t)g * per, R), hi = std::min<uint64_t>(lo + per, R);
            if (hi == lo) break;   // more devices than members
            magpy_b200_ensemble a = *args;
            a.device = devices[g];
            a.n_members = hi - lo;
            a.stream_offset = args->stream_offset + lo;
            if (a.seeds) a.seeds += lo;
            if (a.axis_stride) a.anisotropy_axis += lo * a.axis_stride;
            if (a.m0_stride) a.magnetisation_direction += lo * a.m0_stride;
            if (a.radius_stride) a.radius += lo * a.radius_stride;
            if (a.member_temperature) a.member_temperature += lo;
            if (a.member_index) a.member_index += lo;
            if (a.member_anisotropy) a.member_anisotropy += lo;
            if (a.member_damping) a.member_damping += lo;
            if (a.member_field_amplitude) a.member_field_amplitude += lo;
            if (a.injected_dw) a.injected_dw += lo * a.injected_steps * n;
            if (a.out_trajectories) a.out_trajectories += lo * n * S;
            if (a.out_final) a.out_final += lo * n;
            part.emplace_back(a.out_sums ? S * 4 : 0);
            if (g > 0) a.out_time = a.out_field = nullptr;
            shard.push_back(a);
            magpy_b200_plan* pl = nullptr;
            rc = magpy_b200_plan_create(&shard.back(), &pl);
            if (!rc) plans.push_back(pl);
        }
        // all shards integrate concurrently: plan_run only enqueues work on the plan's own stream
        for (size_t g = 0; g < plans.size() && !rc; ++g) rc = plan_run(plans[g]);
        magpy_b200_stats total;
        std::memset(&total, 0, sizeof total);
        for (size_t g = 0; g < plans.size() && !rc; ++g) {
            const magpy_b200_ensemble& a = shard[g];
            rc = plan_sync(plans[g], nullptr);
            if (!rc)
                rc = plan_fetch(plans[g], a.out_time, a.out_field, a.out_trajectories,
                                a.out_sums ? part[g].data() : nullptr, a.out_final);
            magpy_b200_stats st;
            if (!rc) rc = plan_sync(plans[g], &st);
            if (rc) break;
            total.steps_per_member = st.steps_per_member;
            total.particle_steps += st.particle_steps;
            total.newton_iterations += st.newton_iterations;
            total.newton_max_iterations = std::max(total.newton_max_iterations, st.newton_max_iterations);
            total.newton_failures += st.newton_failures;
            total.kernel_launches += st.kernel_launches;
            total.device_ms = std::max(total.device_ms, st.device_ms);
            total.integrate_ms = std::max(total.integrate_ms, st.integrate_ms);
            total.h2d_bytes += st.h2d_bytes;
            total.d2h_bytes += st.d2h_bytes;
            total.kernel_family = st.kernel_family;
            total.kernel_variant = st.kernel_variant;
        }
        if (!rc && args->out_sums) {   // fixed device order: deterministic for a given device list
            std::fill_n(args->out_sums, S * 4, 0.0);
            for (size_t g = 0; g < plans.size(); ++g)
                for (uint64_t q = 0; q < S * 4; ++q) args->out_sums[q] += part[g][q];
        }
        if (!rc && stats) *stats = total;
    } catch (const std::exception& e) {
        rc = fail(MAGPY_B200_ERR_NOMEM, "%s", e.what());
    }
    for (magpy_b200_plan* pl : plans) delete pl;
    return rc;
}

int magpy_b200_simulate(const double* radius, const double* anisotropy, const double* anisotropy_axis,
                        const double* magnetisation_direction, const double* location, size_t n_particles,
                        double magnetisation, double damping, double temperature, int renorm, int interactions,
                        int use_implicit, double eps, double time_step, double end_time, size_t max_samples,
                        int64_t seed, int field_shape, double field_amplitude, double field_frequency,
                        double* out_time, double* out_field, double* out_m, magpy_b200_stats* stats) {
    magpy_b200_ensemble a;
    std::memset(&a, 0, sizeof a);
    a.abi_version = MAGPY_B200_ABI_VERSION;
    a.device = 0;
    a.n_members = 1;
    a.n_particles = (uint32_t)n_particles;
    a.radius = radius;
    a.anisotropy = anisotropy;
    a.location = location;
    a.anisotropy_axis = anisotropy_axis;
    a.magnetisation_direction = magnetisation_direction;
    a.magnetisation = magnetisation;
    a.damping = damping;
    a.temperature = temperature;
    a.renorm = renorm;
    a.interactions = interactions;
    a.use_implicit = use_implicit;
    a.implicit_tol = eps;
    a.time_step = time_step;
    a.end_time = end_time;
    a.max_samples = max_samples;
    a.field_shape = field_shape;
    a.field_amplitude = field_amplitude;
    a.field_frequency = field_frequency;
    a.seeds = &seed;
    a.gauss_mode = MAGPY_B200_GAUSS_F32_PACKED;
    a.out_time = out_time;
    a.out_field = out_field;
    a.out_trajectories = out_m;  // [1][N][3][S]
    if (n_particles == 0 || n_particles > 0xFFFFFFFFull) return fail(MAGPY_B200_ERR_BAD_ARG, "n_particles out of range");
    if (!out_m) return fail(MAGPY_B200_ERR_BAD_ARG, "out_m is NULL");
    return magpy_b200_simulate_ensemble(&a, stats);
}

int magpy_b200_philox_words(int device, const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    int rc = select_device(device);
    if (rc) return rc;
    DevBuf<uint32_t> d;
    CU_TRY(d.alloc(4));
    CU_TRY(mb::launch_philox_words(ctr, key, d.p));
    CU_TRY(cudaMemcpy(out, d.p, 16, cudaMemcpyDeviceToHost));
    return MAGPY_B200_OK;
}

int magpy_b200_gaussians(int device, int64_t seed, uint64_t member, uint32_t particle, uint64_t first_step,
                         uint64_t n_steps, int gauss_mode, double* out) {
    int rc = select_device(device);
    if (rc) return rc;
    if (!out || n_steps == 0) return fail(MAGPY_B200_ERR_BAD_ARG, "bad arguments");
    DevBuf<double> d;
    CU_TRY(d.alloc(3 * n_steps));
    const int noise = gauss_mode == MAGPY_B200_GAUSS_F64 ? mb::NOISE_PHILOX_F64
                      : gauss_mode == MAGPY_B200_GAUSS_F32 ? mb::NOISE_PHILOX_F32 : mb::NOISE_PHILOX_PACKED;
    CU_TRY(mb::launch_gaussians(noise, (uint64_t)seed, (uint32_t)member, particle, first_step, n_steps, d.p));
    CU_TRY(cudaMemcpy(out, d.p, 3 * n_steps * 8, cudaMemcpyDeviceToHost));
    return MAGPY_B200_OK;
}

int magpy_b200_solve3(int device, size_t n, const double* A, const double* b, double* x, int* ok) {
    int rc = select_device(device);
    if (rc) return rc;
    if (!A || !b || !x || !ok || n == 0) return fail(MAGPY_B200_ERR_BAD_ARG, "bad arguments");
    DevBuf<double> dA, db, dx;
    DevBuf<int> dok;
    CU_TRY(dA.alloc(9 * n));
    CU_TRY(db.alloc(3 * n));
    CU_TRY(dx.alloc(3 * n));
    CU_TRY(dok.alloc(n));
    CU_TRY(cudaMemcpy(dA.p, A, 9 * n * 8, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(db.p, b, 3 * n * 8, cudaMemcpyHostToDevice));
    CU_TRY(mb::launch_solve3(dA.p, db.p, dx.p, dok.p, n));
    CU_TRY(cudaMemcpy(x, dx.p, 3 * n * 8, cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(ok, dok.p, n * sizeof(int), cudaMemcpyDeviceToHost));
    return MAGPY_B200_OK;
}

int magpy_b200_gaussian_stats(int device, int64_t seed, uint64_t first_member, uint64_t n_members, uint64_t n_steps,
                              int gauss_mode, uint64_t* hist, uint64_t* angle_hist, double* moments) {
    int rc = select_device(device);
    if (rc) return rc;
    if (!hist || !angle_hist || !moments || n_members == 0 || n_steps == 0) return fail(MAGPY_B200_ERR_BAD_ARG, "bad arguments");
    if (n_members > 0x7FFFFFFFull) return fail(MAGPY_B200_ERR_BAD_ARG, "n_members too large");
    DevBuf<unsigned long long> d_h, d_a;
    DevBuf<double> d_m;
    CU_TRY(d_h.alloc(4096));
    CU_TRY(d_a.alloc(1024));
    CU_TRY(d_m.alloc(5));
    CU_TRY(cudaMemset(d_h.p, 0, 4096 * 8));
    CU_TRY(cudaMemset(d_a.p, 0, 1024 * 8));
    CU_TRY(cudaMemset(d_m.p, 0, 5 * 8));
    const int noise = gauss_mode == MAGPY_B200_GAUSS_F64 ? mb::NOISE_PHILOX_F64
                      : gauss_mode == MAGPY_B200_GAUSS_F32 ? mb::NOISE_PHILOX_F32 : mb::NOISE_PHILOX_PACKED;
    CU_TRY(mb::launch_gauss_stats(noise, (uint64_t)seed, first_member, n_members, n_steps, d_h.p, d_a.p, d_m.p));
    CU_TRY(cudaMemcpy(hist, d_h.p, 4096 * 8, cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(angle_hist, d_a.p, 1024 * 8, cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(moments, d_m.p, 5 * 8, cudaMemcpyDeviceToHost));
    return MAGPY_B200_OK;
}

int magpy_b200_fp64_peak(int device, double* tflops, double* sm_clock_mhz) {
    int rc = select_device(device);
    if (rc) return rc;
    if (!tflops) return fail(MAGPY_B200_ERR_BAD_ARG, "tflops is NULL");
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    const int iters = 8192;
    cudaEvent_t e0, e1;
    CU_TRY(cudaEventCreate(&e0));
    CU_TRY(cudaEventCreate(&e1));
    double best = 0.0;
    // best over a few launch shapes (resident warps per scheduler) and repetitions
    const int shapes[3][2] = {{128, 8}, {128, 4}, {256, 8}};   // threads per CTA, CTAs per SM
    for (const auto& sh : shapes) {
        const int threads = sh[0], blocks = prop.multiProcessorCount * sh[1];
        DevBuf<double> d;
        CU_TRY(d.alloc((size_t)blocks * threads));
        for (int rep = 0; rep < 4; ++rep) {
            CU_TRY(cudaEventRecord(e0));
            CU_TRY(mb::launch_fp64_peak(d.p, blocks, threads, iters, nullptr));
            CU_TRY(cudaEventRecord(e1));
            CU_TRY(cudaEventSynchronize(e1));
            float ms = 0.f;
            CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
            const double flops = (double)blocks * threads * iters * 64.0 * 2.0;
            if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    if (sm_clock_mhz) {
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
        *sm_clock_mhz = khz / 1000.0;
    }
    return MAGPY_B200_OK;
}

int magpy_b200_simulate_dom(int device, size_t n_items, const double* volume, const double* anisotropy,
                            const double* initial_probabilities, double temperature, double magnetisation, double damping,
                            double time_step, double end_time, size_t max_samples, int field_shape, double field_amplitude,
                            double field_frequency, size_t field_n_components, double* out_time, double* out_field,
                            double* out_mz, uint64_t* out_steps) {
    if (n_items == 0) return fail(MAGPY_B200_ERR_BAD_ARG, "n_items must be >= 1");
    if (!volume || !anisotropy || !initial_probabilities) return fail(MAGPY_B200_ERR_BAD_ARG, "volume/anisotropy/initial_probabilities must not be NULL");
    if (!out_mz) return fail(MAGPY_B200_ERR_BAD_ARG, "out_mz must not be NULL");
    if (max_samples < 2) return fail(MAGPY_B200_ERR_BAD_ARG, "max_samples must be >= 2 (lib/simulation.cpp:703)");
    if (max_samples > 0x7FFFFFFFull) return fail(MAGPY_B200_ERR_BAD_ARG, "max_samples too large");
    if (!(time_step > 0.0) || !(end_time > 0.0)) return fail(MAGPY_B200_ERR_BAD_ARG, "time_step and end_time must be > 0");
    if (field_shape < MAGPY_B200_FIELD_SINE || field_shape > MAGPY_B200_FIELD_SQUARE_FOURIER)
        return fail(MAGPY_B200_ERR_BAD_ARG, "field_shape must be a MAGPY_B200_FIELD_* value");
    if (!(magnetisation > 0.0) || !(temperature > 0.0) || !(damping > 0.0))
        return fail(MAGPY_B200_ERR_BAD_ARG, "magnetisation, temperature and damping must be > 0");
    int rc = select_device(device);
    if (rc) return rc;
    rc = enable_pool(device);
    if (rc) return rc;
    const size_t S = max_samples, n = n_items;
    cudaStream_t stream = nullptr;
    CU_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    struct StreamGuard {
        cudaStream_t s;
        ~StreamGuard() { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
    } guard{stream};
    DevBuf<double> d_vol, d_k, d_p0, d_field, d_mz;
    DevBuf<unsigned long long> d_steps;
    CU_TRY(d_vol.alloc(n, stream));
    CU_TRY(d_k.alloc(n, stream));
    CU_TRY(d_p0.alloc(2 * n, stream));
    CU_TRY(d_field.alloc(n * S, stream));
    CU_TRY(d_mz.alloc(n * S, stream));
    CU_TRY(d_steps.alloc(n, stream));
    CU_TRY(cudaMemcpyAsync(d_vol.p, volume, n * 8, cudaMemcpyHostToDevice, stream));
    CU_TRY(cudaMemcpyAsync(d_k.p, anisotropy, n * 8, cudaMemcpyHostToDevice, stream));
    CU_TRY(cudaMemcpyAsync(d_p0.p, initial_probabilities, 2 * n * 8, cudaMemcpyHostToDevice, stream));
    mb::DomBatch B{};
    B.n = n; B.S = S;
    B.volume = d_vol.p; B.anisotropy = d_k.p; B.p0 = d_p0.p;
    B.temperature = temperature; B.magnetisation = magnetisation; B.alpha = damping; B.mu0 = kMU0;
    B.time_step = time_step; B.end_time = end_time;
    B.field_shape = field_shape; B.field_amplitude = field_amplitude; B.field_frequency = field_frequency;
    B.n_components = (unsigned)field_n_components;
    B.out_field = d_field.p; B.out_mz = d_mz.p; B.out_steps = d_steps.p;
    CU_TRY(mb::launch_dom(B, stream));
    CU_TRY(cudaMemcpyAsync(out_mz, d_mz.p, n * S * 8, cudaMemcpyDeviceToHost, stream));
    if (out_field) CU_TRY(cudaMemcpyAsync(out_field, d_field.p, n * S * 8, cudaMemcpyDeviceToHost, stream));
    if (out_steps) CU_TRY(cudaMemcpyAsync(out_steps, d_steps.p, n * 8, cudaMemcpyDeviceToHost, stream));
    CU_TRY(cudaStreamSynchronize(stream));
    if (out_time) {   // lib/simulation.cpp:704,759-760
        const double sampling_time = end_time / (S - 1);
        out_time[0] = 0;
        for (unsigned int sample = 1; sample < S; sample++) out_time[sample] = sample * sampling_time;
    }
    return MAGPY_B200_OK;
}

int magpy_b200_fp64_mma_peak(int device, double* tflops) {
    int rc = select_device(device);
    if (rc) return rc;
    if (!tflops) return fail(MAGPY_B200_ERR_BAD_ARG, "tflops is NULL");
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    const int iters = 4000, threads = 256, blocks = prop.multiProcessorCount;   // 2 warps per scheduler
    cudaEvent_t e0, e1;
    CU_TRY(cudaEventCreate(&e0));
    CU_TRY(cudaEventCreate(&e1));
    DevBuf<double> d;
    CU_TRY(d.alloc((size_t)blocks * threads));
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        CU_TRY(cudaEventRecord(e0));
        CU_TRY(mb::launch_fp64_mma_peak(d.p, blocks, threads, iters, nullptr));
        CU_TRY(cudaEventRecord(e1));
        CU_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = (double)blocks * (threads / 32) * iters * 48.0 * 512.0;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    return MAGPY_B200_OK;
}

}  // extern "C"
