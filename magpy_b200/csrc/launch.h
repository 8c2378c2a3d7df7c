// launch.h — host-callable launchers of the sm_100a kernels (one translation unit per kernel family,
// compiled in parallel; the host side in magpy_b200.cu sees only these functions).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mb {

struct RunParams;

constexpr int K1_LATENCY = 100;

// noise: NOISE_PHILOX_F32 | NOISE_PHILOX_F64 | NOISE_INJECTED | NOISE_PHILOX_PACKED; tab: applied field from field_tab
// min_blocks: variant of the production (packed-noise) instantiation: 1 (free registers), 7 (7 CTAs per SM) or
// K1_LATENCY (small ensembles: applied-field table entries fetched one step pair ahead)
cudaError_t launch_heun_single(int noise, bool tab, bool axis_z, int min_blocks, unsigned grid, cudaStream_t s,
                               const RunParams& P);
// K1b (heun_single_balanced.cu): the same integration as a persistent kernel over (time segment, member block) tasks
cudaError_t launch_heun_single_balanced(bool tab, bool axis_z, unsigned phys_grid, cudaStream_t s, const RunParams& P);
int heun_single_balanced_resident_ctas(bool tab, bool axis_z, bool renorm);
// K1s (heun_single_split.cu): small ensembles, grid = ceil(R / 32) CTAs of one integrator warp and one generator warp
cudaError_t launch_heun_single_split(bool tab, bool axis_z, int producers, unsigned grid, cudaStream_t s, const RunParams& P);
constexpr int K1_SPLIT = 300;   // stats.kernel_variant of K1s
int heun_single_resident_ctas(bool tab, bool axis_z, bool renorm, int min_blocks);
cudaError_t launch_imid_single(int noise, bool tab, bool axis_z, unsigned grid, cudaStream_t s, const RunParams& P);
cudaError_t launch_heun_small(int noise, bool tab, unsigned n_particles, unsigned grid, cudaStream_t s, const RunParams& P);
cudaError_t launch_imid_small(int noise, bool tab, unsigned n_particles, unsigned grid, cudaStream_t s, const RunParams& P);
cudaError_t launch_imid_split(int noise, bool tab, unsigned n_particles, unsigned grid, cudaStream_t s, const RunParams& P);
// one warp per particle (N = 2..4, small ensembles): grid = ceil(R / 32) CTAs of 32 N threads
cudaError_t launch_imid_warps(int noise, bool tab, unsigned n_particles, unsigned grid, cudaStream_t s, const RunParams& P);
// layout: 0 = pair table in global memory, 1 = table in shared memory, 2 = table in shared memory + one moment buffer
cudaError_t launch_heun_cluster(int noise, bool tab, int np, int layout, dim3 grid, dim3 block, size_t smem,
                                cudaStream_t s, const RunParams& P);
// K2m (cluster_mma.cu): threads = 32 * G * member halves; one_buf = predictor overwrites the current moments
// two launches: `full_ctas` CTAs with a full set of members (whole waves), then `tail_ctas` CTAs of the partial last wave
cudaError_t launch_heun_cluster_mma(int noise, bool tab, bool one_buf, unsigned full_ctas, unsigned tail_ctas, unsigned threads,
                                    size_t smem, cudaStream_t s, const RunParams& P);
// cluster_big.cu: Heun for any cluster size (moments in global memory); CTA = 32 members x 16 particle slots
cudaError_t launch_heun_cluster_big(int noise, bool tab, unsigned grid, cudaStream_t s, const RunParams& P);
// the same layout for the implicit midpoint scheme (two iterate buffers in global memory)
cudaError_t launch_imid_cluster_big(int noise, bool tab, unsigned grid, cudaStream_t s, const RunParams& P);
// K4m (cluster_mma_imid.cu): threads = 32 * G * column tiles of 8 members
cudaError_t launch_imid_cluster_mma(int noise, bool tab, unsigned grid, unsigned threads, size_t smem, cudaStream_t s,
                                    const RunParams& P);
cudaError_t launch_imid_cluster(int noise, bool tab, int np, dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                const RunParams& P);

cudaError_t launch_field_table(double* tab, uint64_t j0, uint64_t n_steps, double dt, double second_offset, int shape,
                               double h0, double f_red, cudaStream_t s);
cudaError_t launch_reduce_partials(const double* partial, double* sums, uint32_t k0, uint32_t n_samples, uint32_t n_cta,
                                   cudaStream_t s);
cudaError_t launch_transpose(const double* in, double* out, uint64_t rows, uint64_t cols, uint64_t batches, uint64_t in_bs,
                             uint64_t in_rs, uint64_t out_bs, uint64_t out_rs, double scale, cudaStream_t s);
// in[S][n][R] -> out[R][n][S] * scale (trajectory fetch)
cudaError_t launch_traj_fetch(const double* in, double* out, uint64_t R, uint32_t n, uint32_t S, double scale,
                              cudaStream_t s);
cudaError_t launch_broadcast_rows(const double* in, double* out, uint64_t n, uint64_t R, cudaStream_t s);
cudaError_t launch_fp64_peak(double* out, int blocks, int threads, int iters, cudaStream_t s);
cudaError_t launch_fp64_mma_peak(double* out, int blocks, int threads, int iters, cudaStream_t s);
cudaError_t launch_philox_words(const uint32_t ctr[4], const uint32_t key[2], uint32_t* out);
cudaError_t launch_gaussians(int noise, uint64_t seed, uint32_t member, uint32_t particle, uint64_t first_step,
                             uint64_t n_steps, double* out);

cudaError_t launch_solve3(const double* A, const double* b, double* x, int* ok, uint64_t n);
// histogram [4096] of the draws over [-8, 8), histogram [1024] of the Box-Muller pair angles, {sum z, z^2, z^3, z^4, max |z|}
cudaError_t launch_gauss_stats(int noise, uint64_t seed, uint64_t first_member, uint64_t n_members, uint64_t n_steps,
                               unsigned long long* hist, unsigned long long* angle_hist, double* moments);

// discrete-orientation model (dom.cu): one thread per batch item
struct DomBatch {
    uint64_t n, S;
    const double* volume;       // [n]
    const double* anisotropy;   // [n]
    const double* p0;           // [n][2]
    double temperature, magnetisation, alpha, mu0;
    double time_step, end_time;
    int field_shape;            // 0 sine, 1 square, 2 constant, 3 square_fourier
    double field_amplitude, field_frequency;
    unsigned n_components;
    double* out_field;          // [n][S]  (A/m)
    double* out_mz;             // [n][S]  (p_0 - p_1)
    unsigned long long* out_steps;   // [n] accepted RK45 steps, or nullptr
};
cudaError_t launch_dom(const DomBatch& B, cudaStream_t s);

}  // namespace mb
