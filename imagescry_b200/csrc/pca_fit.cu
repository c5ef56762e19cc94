// PCA.fit on the GPU: the data-dependent part of
//   /root/reference/src/imagescry/models/decomposition.py:94-148
// i.e. the feature means (:116), the centring (:119) and the second moments of the centred data.
// The reference takes a full SVD of the n x F centred matrix (an n x n `U` is built and thrown away);
// the right singular vectors and singular values it keeps are the eigenvectors / eigenvalues of the
// F x F covariance  C = Xc^T Xc / (n - 1),  which is all the component selection (:125-146) needs.
// The F x F eigen-decomposition (F = 1280: 6.5 MB) is a small dense library call on the host side.
//
//   K6a col_sum_kernel        column sums, fp64 accumulation, row chunks in parallel
//   K6b col_mean_kernel       fixed-order fold of the chunk sums -> fp32 means (correctly rounded)
//   K6c center_split_kernel   per row chunk: Xc = x - mean in fp32 exactly as the reference does, split
//                             into bf16 hi + lo (|Xc - hi - lo| <= 2^-17 |Xc|) and TRANSPOSED to F x rows
//                             (contraction dim contiguous: K-major operand tiles for the tensor cores)
//   K6d cov_syrk_kernel       tcgen05 SYRK: 128 x 256 tiles of Xc^T Xc that touch the upper triangle,
//                             split-K over the chunk's rows; operands by TMA (128B swizzle), three MMAs
//                             per k-step (hi.hi + hi.lo + lo.hi: fp32-class products) into one fp32 TMEM
//                             accumulator pair (hi.hi in one, the small cross terms in the other); every
//                             (tile, split) owns a slot of the partial buffer and adds to it chunk after
//                             chunk (stream-ordered, no atomics: deterministic).  The tensor core's fp32
//                             accumulation TRUNCATES (measured: -2.3e-5 relative after 750 accumulating
//                             MMAs of same-sign products), so one accumulation covers at most 1024 rows
//                             (64 hi.hi MMAs, bias < 2e-6) before it is drained with round-to-nearest adds
//   K6e cov_finalize_kernel   fold the split partials in fp64 in a fixed order, divide by n - 1, mirror
#include "common.cuh"

#include <algorithm>

namespace isx {
namespace {

constexpr int kFitThreads = 256;
constexpr int kMaxChunks = 64;  // row chunks of the column sums
constexpr int SM_ = 128;        // covariance tile rows (TMEM lanes)
constexpr int SN_ = 256;        // covariance tile columns (TMEM columns)
constexpr int SK_ = 64;         // rows of x per k-block: 128 bytes of bf16 = one swizzle row
constexpr int kSyrkStages = 2;
constexpr int kSyrkSplits = 5;  // 30 upper-triangle tiles at F = 1280 x 5 = 150 CTAs ~ one wave of 148 SMs
constexpr uint32_t kSyrkAPart = SM_ * SK_ * 2;  // 16 KB (hi or lo)
constexpr uint32_t kSyrkBPart = SN_ * SK_ * 2;  // 32 KB
constexpr uint32_t kSyrkStageBytes = 2 * kSyrkAPart + 2 * kSyrkBPart;  // 96 KB
constexpr uint32_t kSyrkBarOff = kSyrkStages * kSyrkStageBytes;
constexpr uint32_t kSyrkSmem = kSyrkBarOff + 64 + 1024;
constexpr int kSyrkRowsPerAcc = 1024;  // rows one TMEM accumulation covers (truncation bias, see the header)

struct FitPlan {
  int chunks;                // column-sum row chunks
  long long rows_per_chunk;
  int f_pad;                 // F rounded up to 256
  long long nc;              // rows per SYRK chunk (multiple of 64)
  int mt, nt, tiles;         // 128-row / 256-column tiles, tiles touching the upper triangle
  size_t sums_off, xt_hi_off, xt_lo_off, part_off, total;
};

FitPlan fit_plan(long long n, int F) {
  FitPlan p;
  p.chunks = static_cast<int>(std::max<long long>(1, std::min<long long>(kMaxChunks, (n + 1023) / 1024)));
  p.rows_per_chunk = (n + p.chunks - 1) / p.chunks;
  p.f_pad = (F + SN_ - 1) / SN_ * SN_;
  const long long n_pad = (n + SK_ - 1) / SK_ * SK_;
  p.nc = std::min<long long>(n_pad, static_cast<long long>(kSyrkSplits) * kSyrkRowsPerAcc);
  p.mt = p.f_pad / SM_;
  p.nt = p.f_pad / SN_;
  p.tiles = 0;
  for (int nj = 0; nj < p.nt; ++nj) p.tiles += std::min(p.mt, 2 * nj + 2);
  size_t off = 0;
  p.sums_off = off;
  off += (static_cast<size_t>(p.chunks) * F * sizeof(double) + 255) / 256 * 256;
  p.xt_hi_off = off;
  off += (static_cast<size_t>(p.f_pad) * p.nc * 2 + 1023) / 1024 * 1024;
  p.xt_lo_off = off;
  off += (static_cast<size_t>(p.f_pad) * p.nc * 2 + 1023) / 1024 * 1024;
  p.part_off = off;
  off += static_cast<size_t>(kSyrkSplits) * p.f_pad * p.f_pad * sizeof(float);
  p.total = off + 1024;
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

// K6c: rows [r0, r0 + nc) of x -> xt_hi / xt_lo [f_pad][nc] bf16 (row pitch nc).  32 x 32 tiles through
// shared memory: reads coalesced along the features, writes along the rows.  Rows beyond n and
// features beyond F are written as zeros (they pad the contraction and the tiles).
__global__ void __launch_bounds__(256)
center_split_kernel(const float* __restrict__ x, const float* __restrict__ mean, long long n, int F, long long r0,
                    long long nc, int f_pad, __nv_bfloat16* __restrict__ xt_hi, __nv_bfloat16* __restrict__ xt_lo) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const long long rt = blockIdx.x;                          // row tile inside the chunk
  const int ft = blockIdx.y;
  const int f = ft * 32 + tx;
  const float m = (f < F) ? mean[f] : 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long r = r0 + rt * 32 + ty + 8 * j;
    // centred in fp32 like the reference's `x - self.feature_means` (decomposition.py:119)
    tile[ty + 8 * j][tx] = (r < n && f < F) ? __fsub_rn(x[r * F + f], m) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int fo = ft * 32 + ty + 8 * j;
    const long long c = rt * 32 + tx;
    if (fo < f_pad && c < nc) {
      const float v = tile[tx][ty + 8 * j];
      const __nv_bfloat16 hi = __float2bfloat16_rn(v);
      const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
      xt_hi[static_cast<size_t>(fo) * nc + c] = hi;
      xt_lo[static_cast<size_t>(fo) * nc + c] = lo;
    }
  }
}

// K6d.  grid = (tiles, splits).  warp 0 TMA, warp 1 MMA issuer, warp 2 TMEM alloc, warps 4-7 epilogue.
__global__ void __launch_bounds__(256, 1)
cov_syrk_kernel(const __grid_constant__ CUtensorMap tmap_hi, const __grid_constant__ CUtensorMap tmap_lo, int mt,
                int f_pad, int num_kb, int accumulate, float* __restrict__ part) {
  extern __shared__ uint8_t syrk_raw[];
  uint8_t* smem = syrk_raw + ((1024u - (smem_u32(syrk_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kSyrkBarOff);
  uint64_t* empty_bar = full_bar + kSyrkStages;
  uint64_t* acc_bar = empty_bar + kSyrkStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // tile decode: column block nj holds row blocks mi = 0 .. min(mt, 2 nj + 2) - 1
  int t = blockIdx.x, nj = 0;
  while (t >= min(mt, 2 * nj + 2)) { t -= min(mt, 2 * nj + 2); ++nj; }
  const int mi = t;
  const int split = blockIdx.y, splits = gridDim.y;
  const int kb0 = static_cast<int>(static_cast<long long>(num_kb) * split / splits);
  const int kb1 = static_cast<int>(static_cast<long long>(num_kb) * (split + 1) / splits);

  if (warp == 0 && lane == 0) { prefetch_tmap(&tmap_hi); prefetch_tmap(&tmap_lo); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kSyrkStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) { tmem_alloc(tmem_ptr, 2 * SN_); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = smem + stage * kSyrkStageBytes;
        mbar_arrive_expect_tx(&full_bar[stage], kSyrkStageBytes);
        tma_load_2d(st, &tmap_hi, &full_bar[stage], kb * SK_, mi * SM_, kEvictLast);
        tma_load_2d(st + kSyrkAPart, &tmap_lo, &full_bar[stage], kb * SK_, mi * SM_, kEvictLast);
        uint8_t* bt = st + 2 * kSyrkAPart;
        tma_load_2d(bt, &tmap_hi, &full_bar[stage], kb * SK_, nj * SN_, kEvictLast);
        tma_load_2d(bt + kSyrkAPart, &tmap_hi, &full_bar[stage], kb * SK_, nj * SN_ + SM_, kEvictLast);
        tma_load_2d(bt + kSyrkBPart, &tmap_lo, &full_bar[stage], kb * SK_, nj * SN_, kEvictLast);
        tma_load_2d(bt + kSyrkBPart + kSyrkAPart, &tmap_lo, &full_bar[stage], kb * SK_, nj * SN_ + SM_, kEvictLast);
        if (++stage == kSyrkStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(/*bf16*/ 1, SM_, SN_);
      uint32_t stage = 0, phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(smem + stage * kSyrkStageBytes);
        const uint32_t a_lo = a_hi + kSyrkAPart;
        const uint32_t b_hi = a_hi + 2 * kSyrkAPart;
        const uint32_t b_lo = b_hi + kSyrkBPart;
#pragma unroll
        for (int k = 0; k < SK_ / 16; ++k) {
          const uint32_t o = k * 32;
          const uint64_t dah = make_kmajor_sw128_desc(a_hi + o), dal = make_kmajor_sw128_desc(a_lo + o);
          const uint64_t dbh = make_kmajor_sw128_desc(b_hi + o), dbl = make_kmajor_sw128_desc(b_lo + o);
          const uint32_t first = (kb != kb0 || k != 0) ? 1u : 0u;
          tc_mma_f16(tmem_base, dah, dbh, idesc, first);        // columns [0, 256): hi . hi
          tc_mma_f16(tmem_base + SN_, dah, dbl, idesc, first);  // columns [256, 512): the cross terms
          tc_mma_f16(tmem_base + SN_, dal, dbh, idesc, 1);
        }
        tc_commit(&empty_bar[stage]);
        if (++stage == kSyrkStages) { stage = 0; phase ^= 1; }
      }
      tc_commit(acc_bar);
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;
    const int row = mi * SM_ + ew * 32 + lane;
    float* dst = part + (static_cast<size_t>(split) * f_pad + row) * f_pad + nj * SN_;
    if (kb1 > kb0) {
      mbar_wait(acc_bar, 0);
      tc_fence_after();
    }
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < SN_; c0 += 32) {
      uint32_t r[32], r2[32];
      if (kb1 > kb0) {
        tmem_ld_32x32(taddr + c0, r);
        tmem_ld_32x32(taddr + SN_ + c0, r2);
        tc_wait_ld_regs(r);
        tc_wait_ld_regs(r2);
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                               __uint_as_float(r[j + 3]));
        float4* d4 = reinterpret_cast<float4*>(dst + c0 + j);
        if (accumulate) {
          const float4 o = *d4;
          v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
        }
        *d4 = v;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * SN_);
  }
}

// K6e: cov[i][j] = cov[j][i] = (sum over the split partials, fp64, fixed order) / (n - 1), i <= j
__global__ void cov_finalize_kernel(const float* __restrict__ part, int splits, int F, int f_pad, double denom,
                                    float* __restrict__ cov) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(F) * F) return;
  const int i = static_cast<int>(idx / F), j = static_cast<int>(idx - static_cast<long long>(i) * F);
  if (i > j) return;
  double acc = 0.0;
  for (int s2 = 0; s2 < splits; ++s2) acc += static_cast<double>(part[(static_cast<size_t>(s2) * f_pad + i) * f_pad + j]);
  const float v = static_cast<float>(acc / denom);
  cov[idx] = v;
  cov[static_cast<size_t>(j) * F + i] = v;
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
  __nv_bfloat16* xt_hi = reinterpret_cast<__nv_bfloat16*>(ws + p.xt_hi_off);
  __nv_bfloat16* xt_lo = reinterpret_cast<__nv_bfloat16*>(ws + p.xt_lo_off);
  col_sum_kernel<<<dim3((F + kFitThreads - 1) / kFitThreads, p.chunks), kFitThreads, 0, stream>>>(x, n, F, p.rows_per_chunk, sums);
  ISX_CHECK_CUDA(cudaGetLastError());
  col_mean_kernel<<<(F + 127) / 128, 128, 0, stream>>>(sums, p.chunks, F, static_cast<double>(n), mean);
  ISX_CHECK_CUDA(cudaGetLastError());
  CUtensorMap thi, tlo;
  int rc = encode_tmap_2d(&thi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, xt_hi, static_cast<uint64_t>(p.f_pad),
                          static_cast<uint64_t>(p.nc), static_cast<uint64_t>(p.nc) * 2, SM_, SK_, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != ISX_OK) return rc;
  rc = encode_tmap_2d(&tlo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, xt_lo, static_cast<uint64_t>(p.f_pad),
                      static_cast<uint64_t>(p.nc), static_cast<uint64_t>(p.nc) * 2, SM_, SK_, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != ISX_OK) return rc;
  ISX_CHECK_CUDA(cudaFuncSetAttribute(cov_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSyrkSmem)));
  int chunk_no = 0;
  for (long long r0 = 0; r0 < n; r0 += p.nc, ++chunk_no) {
    const long long rows = std::min<long long>(p.nc, (n - r0 + SK_ - 1) / SK_ * SK_);  // padded to whole k-blocks
    center_split_kernel<<<dim3(static_cast<unsigned>((rows + 31) / 32), p.f_pad / 32), 256, 0, stream>>>(
        x, mean, n, F, r0, p.nc, p.f_pad, xt_hi, xt_lo);
    ISX_CHECK_CUDA(cudaGetLastError());
    cov_syrk_kernel<<<dim3(p.tiles, kSyrkSplits), 256, kSyrkSmem, stream>>>(thi, tlo, p.mt, p.f_pad,
                                                                           static_cast<int>(rows / SK_), chunk_no > 0, part);
    ISX_CHECK_CUDA(cudaGetLastError());
  }
  const long long ff = static_cast<long long>(F) * F;
  cov_finalize_kernel<<<static_cast<int>((ff + 255) / 256), 256, 0, stream>>>(part, kSyrkSplits, F, p.f_pad,
                                                                               static_cast<double>(n - 1), cov);
  ISX_CHECK_CUDA(cudaGetLastError());
  return ISX_OK;
}

}  // extern "C"
