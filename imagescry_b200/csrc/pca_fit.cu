// PCA.fit on the GPU: the data-dependent part of
//   /root/reference/src/imagescry/models/decomposition.py:94-148
// i.e. the feature means (:116), the centring (:119) and the second moments of the centred data.
// The reference takes a full SVD of the n x F centred matrix (an n x n `U` is built and thrown away);
// the right singular vectors and singular values it keeps are the eigenvectors / eigenvalues of the
// F x F covariance  C = Xc^T Xc / (n - 1),  which is all the component selection (:125-146) needs.
// These kernels produce `mean` and `C` in one pass each over the n x F fp32 matrix; the F x F
// eigen-decomposition (F = 1280: 6.5 MB) is a small dense library call on the host side.
//
//   K6a col_sum_kernel       column sums, fp64 accumulation, row chunks in parallel
//   K6b col_mean_kernel      fixed-order fold of the chunk sums -> fp32 means (correctly rounded)
//   K6c cov_partial_kernel   64 x 64 tiles of Xc^T Xc (upper triangle) per row chunk; centring in fp32
//                            exactly as the reference does (x - mean), fp32 FMA accumulation
//   K6d cov_reduce_kernel    fold the chunk partials in fp64, divide by n - 1, mirror
#include "common.cuh"

#include <algorithm>

namespace isx {
namespace {

constexpr int kFitThreads = 256;
constexpr int kTile = 64;       // covariance tile edge
constexpr int kRowStep = 16;    // rows staged per iteration
constexpr int kMaxChunks = 64;  // row chunks

struct FitPlan {
  int chunks;
  long long rows_per_chunk;
  int tiles;  // tiles per edge
  size_t sums_off, part_off, total;
};

FitPlan fit_plan(long long n, int F) {
  FitPlan p;
  p.chunks = static_cast<int>(std::max<long long>(1, std::min<long long>(kMaxChunks, (n + 1023) / 1024)));
  p.rows_per_chunk = (n + p.chunks - 1) / p.chunks;
  p.tiles = (F + kTile - 1) / kTile;
  size_t off = 0;
  p.sums_off = off;
  off += (static_cast<size_t>(p.chunks) * F * sizeof(double) + 255) / 256 * 256;
  p.part_off = off;
  off += static_cast<size_t>(p.chunks) * F * F * sizeof(float);
  p.total = off + 256;
  return p;
}

__global__ void __launch_bounds__(kFitThreads)
col_sum_kernel(const float* __restrict__ x, long long n, int F, long long rows_per_chunk, double* __restrict__ sums) {
  const int c = blockIdx.x * kFitThreads + threadIdx.x;
  const long long r0 = blockIdx.y * rows_per_chunk, r1 = min(n, r0 + rows_per_chunk);
  if (c >= F) return;
  double acc = 0.0;
  for (long long r = r0; r < r1; ++r) acc += static_cast<double>(x[r * F + c]);
  sums[static_cast<size_t>(blockIdx.y) * F + c] = acc;
}

__global__ void col_mean_kernel(const double* __restrict__ sums, int chunks, int F, double n, float* __restrict__ mean) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= F) return;
  double acc = 0.0;
  for (int i = 0; i < chunks; ++i) acc += sums[static_cast<size_t>(i) * F + c];
  mean[c] = static_cast<float>(acc / n);
}

// grid = (upper-triangular tile pairs, chunks).  Thread (ty, tx) of 16 x 16 owns a 4 x 4 block.
__global__ void __launch_bounds__(kFitThreads)
cov_partial_kernel(const float* __restrict__ x, const float* __restrict__ mean, long long n, int F, int tiles,
                   long long rows_per_chunk, float* __restrict__ part) {
  __shared__ float a_s[kRowStep][kTile + 4], b_s[kRowStep][kTile + 4];
  // decode the tile pair (ti <= tj) from a linear index over the upper triangle
  int pair = blockIdx.x, ti = 0;
  while (pair >= tiles - ti) { pair -= tiles - ti; ++ti; }
  const int tj = ti + pair;
  const int i0 = ti * kTile, j0 = tj * kTile;
  const long long r0 = blockIdx.y * rows_per_chunk, r1 = min(n, r0 + rows_per_chunk);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // staging: 256 threads load 16 rows x 64 columns for each operand (4 floats per thread)
  const int lr = threadIdx.x >> 4, lc = (threadIdx.x & 15) * 4;
  for (long long r = r0; r < r1; r += kRowStep) {
    const long long row = r + lr;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int ci = i0 + lc + q, cj = j0 + lc + q;
      // centred in fp32 like the reference's `x - self.feature_means` (decomposition.py:119)
      a_s[lr][lc + q] = (row < r1 && ci < F) ? __fsub_rn(x[row * F + ci], mean[ci]) : 0.f;
      b_s[lr][lc + q] = (row < r1 && cj < F) ? __fsub_rn(x[row * F + cj], mean[cj]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kRowStep; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = a_s[k][ty * 4 + i]; b[i] = b_s[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* dst = part + static_cast<size_t>(blockIdx.y) * F * F;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = i0 + ty * 4 + i, cj = j0 + tx * 4 + j;
      if (ci < F && cj < F) dst[static_cast<size_t>(ci) * F + cj] = acc[i][j];
    }
}

__global__ void cov_reduce_kernel(const float* __restrict__ part, int chunks, int F, double denom, float* __restrict__ cov) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(F) * F) return;
  const int i = static_cast<int>(idx / F), j = static_cast<int>(idx - static_cast<long long>(i) * F);
  if ((i / kTile) > (j / kTile)) return;  // lower tiles are mirrored from the upper ones
  double acc = 0.0;
  for (int c = 0; c < chunks; ++c) acc += static_cast<double>(part[static_cast<size_t>(c) * F * F + idx]);
  const float v = static_cast<float>(acc / denom);
  cov[idx] = v;
  if ((i / kTile) < (j / kTile)) cov[static_cast<size_t>(j) * F + i] = v;
}

// diagonal tiles were computed in full (both triangles): make them exactly symmetric
__global__ void cov_symmetrize_diag_kernel(int F, float* __restrict__ cov) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(F) * F) return;
  const int i = static_cast<int>(idx / F), j = static_cast<int>(idx - static_cast<long long>(i) * F);
  if ((i / kTile) == (j / kTile) && i < j) cov[static_cast<size_t>(j) * F + i] = cov[idx];
}

}  // namespace
}  // namespace isx

using namespace isx;

extern "C" {

size_t isx_pca_moments_workspace_bytes(int64_t n, int F) {
  if (n <= 0 || F <= 0) return 0;
  return fit_plan(n, F).total;
}

int isx_pca_moments(const float* x, int64_t n, int F, float* mean, float* cov, void* workspace, size_t workspace_bytes,
                    isx_stream_t stream_) {
  const char* fn = "isx_pca_moments";
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ISX_REQUIRE(n >= 2 && F > 0, "%s: need at least 2 samples and 1 feature (n=%lld F=%d)", fn, (long long)n, F);
  ISX_REQUIRE(F <= 8192, "%s: at most 8192 features (F=%d)", fn, F);
  ISX_REQUIRE(x && mean && cov, "%s: null pointer", fn);
  const FitPlan p = fit_plan(n, F);
  ISX_REQUIRE(workspace && workspace_bytes >= p.total, "%s: workspace too small (%zu < %zu)", fn, workspace_bytes, p.total);
  uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  double* sums = reinterpret_cast<double*>(ws + p.sums_off);
  float* part = reinterpret_cast<float*>(ws + p.part_off);
  col_sum_kernel<<<dim3((F + kFitThreads - 1) / kFitThreads, p.chunks), kFitThreads, 0, stream>>>(x, n, F, p.rows_per_chunk, sums);
  ISX_CHECK_CUDA(cudaGetLastError());
  col_mean_kernel<<<(F + 127) / 128, 128, 0, stream>>>(sums, p.chunks, F, static_cast<double>(n), mean);
  ISX_CHECK_CUDA(cudaGetLastError());
  const int pairs = p.tiles * (p.tiles + 1) / 2;
  cov_partial_kernel<<<dim3(pairs, p.chunks), kFitThreads, 0, stream>>>(x, mean, n, F, p.tiles, p.rows_per_chunk, part);
  ISX_CHECK_CUDA(cudaGetLastError());
  const long long ff = static_cast<long long>(F) * F;
  const int blocks = static_cast<int>((ff + 255) / 256);
  cov_reduce_kernel<<<blocks, 256, 0, stream>>>(part, p.chunks, F, static_cast<double>(n - 1), cov);
  ISX_CHECK_CUDA(cudaGetLastError());
  cov_symmetrize_diag_kernel<<<blocks, 256, 0, stream>>>(F, cov);
  ISX_CHECK_CUDA(cudaGetLastError());
  return ISX_OK;
}

}  // extern "C"
