// Micro-benchmark: does tcgen05.ld (epilogue reads of one accumulator stage) slow down while
// tcgen05.mma accumulates into the OTHER stage of the same CTA's tensor memory, and vice versa?
// One CTA per SM, cta_group::1, M = 128, N = 256, K = 16, bf16 operands (zeros) from shared memory.
//   warp 0     one lane issues `mmas` MMAs into columns [0, 256) (when mode & 1)
//   warps 4..  each issues `loads` 32x32b.x32 loads from columns [256, 512) (when mode & 2)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I imagescry_b200/csrc -I include \
//        -o tools/_ubench_tmem_mma tools/ubench_tmem_mma.cu -lcuda
#include "common.cuh"

#include <cstdio>

using namespace isx;

__global__ void __launch_bounds__(384, 1)
tmem_mma_kernel(int mode, int mmas, int loads, int epi_warps, int n_cols, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 2) { tmem_alloc(&tmem_ptr, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  unsigned long long cyc = 0;
  if (warp == 0 && lane == 0 && (mode & 1)) {
    const uint32_t idesc = make_idesc(1, 128, static_cast<uint32_t>(n_cols));
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 16384);
    const long long t0 = clock64();
    for (int i = 0; i < mmas; ++i) {
      const int k = i & 3;
      tc_mma_f16(tmem_base, make_kmajor_sw128_desc(a_addr + k * 32), make_kmajor_sw128_desc(b_addr + k * 32), idesc, 1);
    }
    tc_commit(&bar);
    mbar_wait(&bar, 0);
    cyc = clock64() - t0;
    out[blockIdx.x * 2 + 0] = cyc;
  }
  if (warp >= 4 && warp < 4 + epi_warps && (mode & 2)) {
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + 256;
    float acc = 0.f;
    const long long t0 = clock64();
    for (int i = 0; i < loads; ++i) {
      uint32_t r[32];
      tmem_ld_32x32(taddr + ((i * 32) & 255), r);
      tc_wait_ld();
      float m = -1e30f;
#pragma unroll
      for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(r[j]));
      acc += m;
    }
    cyc = clock64() - t0;
    if (lane == 0 && warp == 4) out[blockIdx.x * 2 + 1] = cyc;
    if (acc == 123.f) out[0] = 0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int main() {
  unsigned long long* d;
  cudaMalloc(&d, 148 * 2 * sizeof(unsigned long long));
  cudaFuncSetAttribute(tmem_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768 + 1024);
  const int mmas = 4096, loads = 4096;
  for (int n_cols : {256, 128}) {
    for (int epi_warps : {4, 8}) {
      for (int mode : {1, 2, 3}) {
        unsigned long long h[148 * 2];
        for (int rep = 0; rep < 2; ++rep) {
          cudaMemset(d, 0, sizeof(h));
          tmem_mma_kernel<<<148, 384, 16384 + 32768 + 1024>>>(mode, mmas, loads, epi_warps, n_cols, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        unsigned long long m = 0, l = 0;
        for (int i = 0; i < 148; ++i) { m = h[2 * i] > m ? h[2 * i] : m; l = h[2 * i + 1] > l ? h[2 * i + 1] : l; }
        printf("N=%d epilogue warps %d mode %d (%s): %.1f cycles per MMA, %.1f cycles per 4 KB load per warp\n", n_cols, epi_warps,
               mode, mode == 1 ? "mma only" : mode == 2 ? "ld only" : "both", static_cast<double>(m) / mmas,
               static_cast<double>(l) / loads);
      }
    }
  }
  return 0;
}
