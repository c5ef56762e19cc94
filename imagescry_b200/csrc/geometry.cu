// region-of-interest masks on the feature-map grid and the masked pooling that
// turns an ROI into a query vector for stage 3.
//
//   roi_rasterize_kernel   /root/reference/src/imagescry/geometry.py:14-65 `create_roi_mask`:
//                          rasterio.features.rasterize(shapes, out_shape=(hf, wf),
//                          transform=Affine.scale(w / wf, h / hf), fill=0, all_touched=True) * class_index.
//                          rasterio / GDAL are not part of the reference tree; the restatement is:
//                          a cell is burned when its open rectangle [j sx, (j+1) sx] x [i sy, (i+1) sy]
//                          (image coordinates) shares positive area with a polygon, i.e. when a
//                          polygon edge passes through the open rectangle or the rectangle's centre
//                          lies inside the polygon (even-odd over all its rings).  Cells that only
//                          touch a polygon along an edge or at a corner stay 0, which is what the
//                          reference's tests expect (tests/test_geometry.py:10-52).
//   masked_pool_kernel     no reference code (the annotator app is the consumer): the mean of the
//                          feature-map cells whose mask value equals `class_index`, per image —
//                          B x E x h x w fp32 -> B x E fp32.  HBM-bound: every map is read once.
#include "common.cuh"

#include <algorithm>

namespace isx {
namespace {

// true if the segment p + t d, t in [0, 1], has a point strictly inside (lo, hi) on this axis range
__device__ __forceinline__ bool clip_axis(double p, double d, double lo, double hi, double& t0, double& t1) {
  if (d == 0.0) return p > lo && p < hi;
  double a = (lo - p) / d, b = (hi - p) / d;
  if (a > b) { const double t = a; a = b; b = t; }
  t0 = fmax(t0, a);
  t1 = fmin(t1, b);
  return t0 < t1;
}

__global__ void __launch_bounds__(256)
roi_rasterize_kernel(const float* __restrict__ edges, const int32_t* __restrict__ poly_offsets, int n_poly,
                     double sx, double sy, int fh, int fw, long long class_index, long long* __restrict__ mask) {
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= fh * fw) return;
  const int i = cell / fw, j = cell - i * fw;
  const double x0 = j * sx, x1 = (j + 1) * sx, y0 = i * sy, y1 = (i + 1) * sy;
  const double cx = 0.5 * (x0 + x1), cy = 0.5 * (y0 + y1);
  bool burned = false;
  for (int pgn = 0; pgn < n_poly && !burned; ++pgn) {
    bool inside = false;  // even-odd parity of the centre
    bool crosses = false;
    for (int e = poly_offsets[pgn]; e < poly_offsets[pgn + 1]; ++e) {
      const double ax = edges[4 * e + 0], ay = edges[4 * e + 1], bx = edges[4 * e + 2], by = edges[4 * e + 3];
      double t0 = 0.0, t1 = 1.0;
      if (clip_axis(ax, bx - ax, x0, x1, t0, t1) && clip_axis(ay, by - ay, y0, y1, t0, t1)) { crosses = true; break; }
      // ray from the centre towards +x
      if ((ay > cy) != (by > cy)) {
        const double xi = ax + (cy - ay) * (bx - ax) / (by - ay);
        if (xi > cx) inside = !inside;
      }
    }
    burned = crosses || inside;
  }
  mask[cell] = burned ? class_index : 0;
}

// One warp per (image, feature) row of hw cells; the mask row (1 byte per cell after the compare)
// is staged in shared memory once per CTA.  Blocks walk (image, feature-chunk) pairs.
__global__ void __launch_bounds__(256)
masked_pool_kernel(const float* __restrict__ fmap, int B, int E, int hw, const long long* __restrict__ mask,
                   int mask_per_image, long long class_index, float* __restrict__ out) {
  extern __shared__ uint8_t sel[];  // [hw]
  __shared__ int s_count;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunks = (E + 63) / 64;  // 64 features per block step: 8 warps x 8 rows
  int staged_img = -1;
  for (long long item = blockIdx.x; item < static_cast<long long>(B) * chunks; item += gridDim.x) {
    const int img = static_cast<int>(item / chunks), chunk = static_cast<int>(item - static_cast<long long>(img) * chunks);
    if (staged_img != img && (mask_per_image || staged_img < 0)) {
      __syncthreads();
      if (threadIdx.x == 0) s_count = 0;
      __syncthreads();
      const long long* m = mask + (mask_per_image ? static_cast<long long>(img) * hw : 0);
      int local = 0;
      for (int c = threadIdx.x; c < hw; c += blockDim.x) {
        const uint8_t on = m[c] == class_index ? 1 : 0;
        sel[c] = on;
        local += on;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(kFullMask, local, o);
      if (lane == 0 && local) atomicAdd(&s_count, local);
      __syncthreads();
    }
    staged_img = img;
    const int count = s_count;
    const float inv = count > 0 ? 1.0f / static_cast<float>(count) : 0.f;
    for (int r = warp; r < 64; r += 8) {
      const int f = chunk * 64 + r;
      if (f >= E) break;
      const float* row = fmap + (static_cast<long long>(img) * E + f) * hw;
      float acc = 0.f;
      for (int c = lane; c < hw; c += 32) acc += sel[c] ? row[c] : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFullMask, acc, o);
      if (lane == 0) out[static_cast<long long>(img) * E + f] = acc * inv;
    }
  }
}

}  // namespace
}  // namespace isx

using namespace isx;

extern "C" {

int isx_roi_rasterize(const float* edges, const int32_t* poly_offsets, int n_poly, int image_h, int image_w,
                      int fmap_h, int fmap_w, int64_t class_index, int64_t* mask, isx_stream_t stream_) {
  const char* fn = "isx_roi_rasterize";
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ISX_REQUIRE(n_poly >= 0 && image_h > 0 && image_w > 0 && fmap_h > 0 && fmap_w > 0,
              "%s: need n_poly >= 0 and positive shapes (n_poly=%d image=%dx%d fmap=%dx%d)", fn, n_poly, image_h, image_w,
              fmap_h, fmap_w);
  ISX_REQUIRE(mask && (n_poly == 0 || (edges && poly_offsets)), "%s: null pointer", fn);
  const int cells = fmap_h * fmap_w;
  roi_rasterize_kernel<<<(cells + 255) / 256, 256, 0, stream>>>(
      edges, poly_offsets, n_poly, static_cast<double>(image_w) / fmap_w, static_cast<double>(image_h) / fmap_h, fmap_h,
      fmap_w, static_cast<long long>(class_index), reinterpret_cast<long long*>(mask));
  ISX_CHECK_CUDA(cudaGetLastError());
  return ISX_OK;
}

int isx_masked_pool(const float* fmap, int B, int E, int h, int w, const int64_t* mask, int mask_per_image,
                    int64_t class_index, float* out, isx_stream_t stream_) {
  const char* fn = "isx_masked_pool";
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ISX_REQUIRE(B >= 0 && E > 0 && h > 0 && w > 0, "%s: need B >= 0 and E, h, w > 0 (B=%d E=%d h=%d w=%d)", fn, B, E, h, w);
  ISX_REQUIRE(static_cast<long long>(h) * w <= 48 * 1024, "%s: at most 49152 cells per map (h*w=%lld)", fn,
              static_cast<long long>(h) * w);
  if (B == 0) return ISX_OK;
  ISX_REQUIRE(fmap && mask && out, "%s: null pointer", fn);
  int sms = 148;
  int rc = device_sm_count(&sms);
  if (rc != ISX_OK) return rc;
  const long long items = static_cast<long long>(B) * ((E + 63) / 64);
  const int blocks = static_cast<int>(std::max<long long>(1, std::min<long long>(items, static_cast<long long>(sms) * 8)));
  masked_pool_kernel<<<blocks, 256, static_cast<size_t>(h) * w, stream>>>(
      fmap, B, E, h * w, reinterpret_cast<const long long*>(mask), mask_per_image ? 1 : 0,
      static_cast<long long>(class_index), out);
  ISX_CHECK_CUDA(cudaGetLastError());
  return ISX_OK;
}

}  // extern "C"
