// Micro-benchmark: tcgen05.ld throughput per SM as a function of the number of reading warps and of
// how many loads are kept in flight before tcgen05.wait::ld.  Decides whether the fused top-k
// epilogue of knn_search_kernel at d = 256 (128 KB of accumulator reads per 2048 MMA cycles) is
// TMEM-bandwidth-bound or issue/latency-bound.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_ubench_tmem tools/ubench_tmem.cu
//   ./tools/_ubench_tmem
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// mode 0: ld, wait, consume.  mode 1: two loads in flight (software pipeline).  mode 2: ld only
// (4 back to back, one wait) — the pure bandwidth figure.
template <int MODE>
__global__ void __launch_bounds__(512, 1) tmem_read_kernel(int iters, unsigned long long* cycles, float* sink) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_ptr + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  const unsigned long long t0 = clock64();
  if (MODE == 0) {
    for (int it = 0; it < iters; ++it) {
      uint32_t r[32];
      tmem_ld32(base + ((it * 32) & 511), r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float m = -1e30f;
#pragma unroll
      for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(r[j]) * 1.0001f);
      acc += m;
    }
  } else if (MODE == 1) {
    uint32_t ra[32], rb[32];
    tmem_ld32(base, ra);
    for (int it = 0; it < iters; it += 2) {
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tmem_ld32(base + (((it + 1) * 32) & 511), rb);
      float m = -1e30f;
#pragma unroll
      for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(ra[j]) * 1.0001f);
      acc += m;
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tmem_ld32(base + (((it + 2) * 32) & 511), ra);
      m = -1e30f;
#pragma unroll
      for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(rb[j]) * 1.0001f);
      acc += m;
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    acc += __uint_as_float(ra[0]);
  } else {
    for (int it = 0; it < iters; it += 4) {
      uint32_t r0[32], r1[32], r2[32], r3[32];
      tmem_ld32(base + ((it * 32) & 511), r0);
      tmem_ld32(base + (((it + 1) * 32) & 511), r1);
      tmem_ld32(base + (((it + 2) * 32) & 511), r2);
      tmem_ld32(base + (((it + 3) * 32) & 511), r3);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += __uint_as_float(r0[0] ^ r1[5] ^ r2[9] ^ r3[31]);
    }
  }
  __syncthreads();
  const unsigned long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_ptr), "r"(512u) : "memory");
  }
}

template <int MODE>
void run(int warps, int iters, unsigned long long* d_cyc, float* d_sink) {
  tmem_read_kernel<MODE><<<148, warps * 32>>>(iters, d_cyc, d_sink);
  cudaDeviceSynchronize();
  tmem_read_kernel<MODE><<<148, warps * 32>>>(iters, d_cyc, d_sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return; }
  unsigned long long h[148];
  cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
  unsigned long long mx = 0;
  for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  const double bytes = static_cast<double>(warps) * iters * 4096.0;
  printf("mode %d warps %2d: %8llu cycles, %.1f B/cycle/SM, %.1f cycles per 4 KB load per warp\n", MODE, warps, mx,
         bytes / mx, static_cast<double>(mx) / iters);
}

int main() {
  unsigned long long* d_cyc;
  float* d_sink;
  cudaMalloc(&d_cyc, 148 * sizeof(unsigned long long));
  cudaMalloc(&d_sink, 4);
  const int iters = 4096;
  for (int warps : {1, 2, 4, 8, 16}) {
    run<0>(warps, iters, d_cyc, d_sink);
    run<1>(warps, iters, d_cyc, d_sink);
    run<2>(warps, iters, d_cyc, d_sink);
  }
  return 0;
}
