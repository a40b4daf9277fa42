// k_prep.cu — K0: device-side frame preparation.
//   * the uint8 -> fp32 cast and the central-difference gradient planes of the reference
//     (DescriptorFrame::Create src/photobundle.cc:225-232, imgradient src/imgproc.cc:27-106)
//     are fused into K_A's footprint staging, so no kernel is needed for them;
//   * this file holds the pyramid step the reference intends at src/photobundle_pyramid.cc:46
//     (cv::pyrDown): separable [1 4 6 4 1]/16 blur with BORDER_REFLECT_101, keep the even samples,
//     dst size (n+1)/2 (src/types.h:70-73), uint8 arithmetic = (sum + 128) >> 8 — bit-identical to
//     cv::pyrDown for CV_8U (checked in tests/test_pyramid.py against cv2).
#include "pba_device.cuh"

namespace pba {

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
  return i;
}

// one thread per destination pixel; source rows are pitched
__global__ void __launch_bounds__(256) k_pyrdown_u8(const uint8_t* __restrict__ src, int srows, int scols, int spitch,
                                                    uint8_t* __restrict__ dst, int drows, int dcols, int dpitch) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= dcols || y >= drows) return;
  const int w[5] = {1, 4, 6, 4, 1};
  int xs[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) xs[i] = reflect101(2 * x + i - 2, scols);
  int acc = 0;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const uint8_t* row = src + (size_t)reflect101(2 * y + j - 2, srows) * spitch;
    int h = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) h += w[i] * (int)row[xs[i]];
    acc += w[j] * h;
  }
  dst[(size_t)y * dpitch + x] = (uint8_t)((acc + 128) >> 8);
}

cudaError_t launch_pyrdown_u8(const uint8_t* src, int srows, int scols, int spitch, uint8_t* dst, int dpitch,
                              cudaStream_t stream) {
  const int drows = (srows + 1) / 2, dcols = (scols + 1) / 2;
  dim3 grid((dcols + 31) / 32, (drows + 7) / 8);
  k_pyrdown_u8<<<grid, 256, 0, stream>>>(src, srows, scols, spitch, dst, drows, dcols, dpitch);
  return cudaGetLastError();
}

// ---- multi-channel descriptor construction -------------------------------------------
// DescriptorFrame::Create (src/photobundle.cc:220-248) on the device.  Planes are fp32 [C][rows][pitch].
//   IntensityAndGradient: {I, Ix, Iy} with imgradient's 0.5*(a-b) and zero first/last row and column
//     (src/imgproc.cc:27-106, scale src/imgproc.h:54-58);
//   BitPlanes (computeBitPlanes, src/imgproc.cc:222-245): cv::GaussianBlur(3x3, sigma 1) of the uint8 image
//     [fixed point: kernel {70,116,70}/256 per axis, (sum + 2^15) >> 16, BORDER_REFLECT_101], the 8-neighbour
//     census transform with `>=` (src/imgproc.cc:140-160; zero first/last row and column), then channel b =
//     bit b of the census byte as 0/1 float, blurred by cv::GaussianBlur(5x5, sigma 1.5) in fp32
//     [separable, row pass then column pass, k0*c + k1*(l1+r1) + k2*(l2+r2), BORDER_REFLECT_101].

__global__ void __launch_bounds__(256) k_channels_ig(const uint8_t* __restrict__ src, int rows, int cols, int spitch,
                                                     float* __restrict__ dst, int dpitch, size_t dplane) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= cols || y >= rows) return;
  const uint8_t* r = src + (size_t)y * spitch;
  float gx = 0.f, gy = 0.f;
  if (x > 0 && x < cols - 1 && y > 0 && y < rows - 1) {
    gx = __fmul_rn(0.5f, __fsub_rn((float)r[x + 1], (float)r[x - 1]));
    gy = __fmul_rn(0.5f, __fsub_rn((float)r[x + spitch], (float)r[x - spitch]));
  }
  const size_t o = (size_t)y * dpitch + x;
  dst[o] = (float)r[x];
  dst[dplane + o] = gx;
  dst[2 * dplane + o] = gy;
}

__global__ void __launch_bounds__(256) k_gauss3_u8(const uint8_t* __restrict__ src, int rows, int cols, int spitch,
                                                   uint8_t* __restrict__ dst, int dpitch) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= cols || y >= rows) return;
  const int w[3] = {70, 116, 70};
  int acc = 0;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const uint8_t* row = src + (size_t)reflect101(y + j - 1, rows) * spitch;
    int h = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) h += w[i] * (int)row[reflect101(x + i - 1, cols)];
    acc += w[j] * h;
  }
  dst[(size_t)y * dpitch + x] = (uint8_t)((acc + 32768) >> 16);
}

__global__ void __launch_bounds__(256) k_census(const uint8_t* __restrict__ src, int rows, int cols, int spitch,
                                                uint8_t* __restrict__ dst, int dpitch) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= cols || y >= rows) return;
  unsigned v = 0;
  if (x > 0 && x < cols - 1 && y > 0 && y < rows - 1) {
    const uint8_t* r = src + (size_t)y * spitch + x;
    const unsigned c = r[0];
    v = (r[-spitch - 1] >= c ? 0x01u : 0u) | (r[-spitch] >= c ? 0x02u : 0u) | (r[-spitch + 1] >= c ? 0x04u : 0u) |
        (r[-1] >= c ? 0x08u : 0u) | (r[1] >= c ? 0x10u : 0u) | (r[spitch - 1] >= c ? 0x20u : 0u) |
        (r[spitch] >= c ? 0x40u : 0u) | (r[spitch + 1] >= c ? 0x80u : 0u);
  }
  dst[(size_t)y * dpitch + x] = (uint8_t)v;
}

// fp32 kernel of cv::getGaussianKernel(5, 1.5, CV_32F)
__device__ __constant__ float c_g5[3] = {0x1.2b1778p-2f, 0x1.defcep-3f, 0x1.ebd75p-4f};   // centre, +-1, +-2 (0.29208171, 0.23388076, 0.12007838)

// one thread per pixel, all 8 bit planes: row pass of the five source rows, then the column pass
__global__ void __launch_bounds__(256) k_bitplanes(const uint8_t* __restrict__ census, int rows, int cols, int spitch,
                                                   float* __restrict__ dst, int dpitch, size_t dplane) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= cols || y >= rows) return;
  int xs[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) xs[i] = reflect101(x + i - 2, cols);
  unsigned char px[5][5];
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const uint8_t* row = census + (size_t)reflect101(y + j - 2, rows) * spitch;
#pragma unroll
    for (int i = 0; i < 5; ++i) px[j][i] = row[xs[i]];
  }
  const float k0 = c_g5[0], k1 = c_g5[1], k2 = c_g5[2];
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    float rp[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const float a0 = (float)((px[j][0] >> b) & 1), a1 = (float)((px[j][1] >> b) & 1), a2 = (float)((px[j][2] >> b) & 1),
                  a3 = (float)((px[j][3] >> b) & 1), a4 = (float)((px[j][4] >> b) & 1);
      rp[j] = __fadd_rn(__fadd_rn(__fmul_rn(k0, a2), __fmul_rn(k1, __fadd_rn(a1, a3))), __fmul_rn(k2, __fadd_rn(a0, a4)));
    }
    dst[(size_t)b * dplane + (size_t)y * dpitch + x] =
        __fadd_rn(__fadd_rn(__fmul_rn(k0, rp[2]), __fmul_rn(k1, __fadd_rn(rp[1], rp[3]))), __fmul_rn(k2, __fadd_rn(rp[0], rp[4])));
  }
}

// saliency of a frame = sum over channels of |Ix| + |Iy| (DescriptorFrame::computeSaliencyMap,
// src/photobundle.cc:212-220; gradients by imgradient on the fp32 channel, zero borders)
__global__ void __launch_bounds__(256) k_saliency(const float* __restrict__ planes, int n_channels, int rows, int cols, int pitch,
                                                  size_t plane, float* __restrict__ out /*dense rows x cols*/) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= cols || y >= rows) return;
  float acc = 0.f;
  if (x > 0 && x < cols - 1 && y > 0 && y < rows - 1) {
    for (int k = 0; k < n_channels; ++k) {
      const float* r = planes + (size_t)k * plane + (size_t)y * pitch + x;
      const float gx = __fmul_rn(0.5f, __fsub_rn(r[1], r[-1])), gy = __fmul_rn(0.5f, __fsub_rn(r[pitch], r[-pitch]));
      const float m = __fadd_rn(fabsf(gx), fabsf(gy));
      acc = (k == 0) ? m : __fadd_rn(acc, m);
    }
  }
  out[(size_t)y * cols + x] = acc;
}

// ExtractPatch (src/photobundle.cc:466-479): integer-pixel patch of every channel, coordinates clamped to
// [radius, size - radius - 1]; channel-major, row-major inside a channel; widened to double.
__global__ void __launch_bounds__(128) k_extract_patches(const float* __restrict__ planes, int n_channels, int rows, int cols,
                                                         int pitch, size_t plane, int radius, int n,
                                                         const int* __restrict__ xy, double* __restrict__ desc) {
  const int side = 2 * radius + 1, P = side * side, CP = n_channels * P;
  const int p = blockIdx.x;
  if (p >= n) return;
  const int ux = xy[2 * p], uy = xy[2 * p + 1];
  for (int e = threadIdx.x; e < CP; e += blockDim.x) {
    const int k = e / P, j = e - k * P, r = j / side - radius, c = j % side - radius;
    const int ri = max(radius, min(uy + r, rows - radius - 1)), ci = max(radius, min(ux + c, cols - radius - 1));
    desc[(size_t)p * CP + e] = (double)planes[(size_t)k * plane + (size_t)ri * pitch + ci];
  }
}

// ---- addFrame on the device -----------------------------------------------------------
// interp2 (src/photobundle.cc:262-294) on the uint8 image with the reference's mixed float/double
// promotions spelled out (the host build uses no FMA contraction either).
__device__ __forceinline__ float interp2_u8(const uint8_t* __restrict__ I, int rows, int cols, int pitch, float xf, float yf) {
  const int max_cols = cols - 1, max_rows = rows - 1;
  const int xi = (int)floorf(xf), yi = (int)floorf(yf);
  xf = __fsub_rn(xf, (float)xi); yf = __fsub_rn(yf, (float)yi);
  auto at = [&](int y, int x) -> float { return (float)I[(size_t)y * pitch + x]; };
  if (xi >= 0 && xi < max_cols && yi >= 0 && yi < max_rows) {
    const float wx = __double2float_rn(__dsub_rn(1.0, (double)xf));
    const float t0 = __fadd_rn(__fmul_rn(at(yi, xi), wx), __fmul_rn(at(yi, xi + 1), xf));
    const float t1 = __fadd_rn(__fmul_rn(at(yi + 1, xi), wx), __fmul_rn(at(yi + 1, xi + 1), xf));
    return __double2float_rn(__dadd_rn(__dmul_rn(__dsub_rn(1.0, (double)yf), (double)t0), (double)__fmul_rn(yf, t1)));
  }
  if (xi == max_cols && yi < max_rows && yi >= 0)
    return (xf > 0) ? 0.f : __double2float_rn(__dadd_rn(__dmul_rn(__dsub_rn(1.0, (double)yf), (double)at(yi, xi)), (double)__fmul_rn(yf, at(yi + 1, xi))));
  if (yi == max_rows && xi < max_cols && xi >= 0)
    return (yf > 0) ? 0.f : __double2float_rn(__dadd_rn(__dmul_rn(__dsub_rn(1.0, (double)xf), (double)at(yi, xi)), (double)__fmul_rn(xf, at(yi, xi + 1))));
  if (xi == max_cols && yi == max_rows) return (xf > 0 || yf > 0) ? 0.f : at(yi, xi);
  return 0.f;
}

// Data association of addFrame (src/photobundle.cc:508-542): one thread per live scene point.
//   uv = normHomog(K * (T_c * X)); tested iff the rounded pixel lies inside the border band; score =
//   ZnccPatch_<2,float>::score (:315-361) of the stored (mean-free) patch against the patch interpolated
//   at uv in the new image.  score = -2 marks "not tested".
__global__ void __launch_bounds__(128) k_associate(const uint8_t* __restrict__ img, int rows, int cols, int pitch, int n,
                                                   const double* __restrict__ xyz, const float* __restrict__ ref_patch,
                                                   const float* __restrict__ ref_norm, const double* __restrict__ Tc /*col-major 4x4*/,
                                                   const double* __restrict__ K /*row-major 3x3*/, int border,
                                                   float* __restrict__ score, int* __restrict__ rc) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const double x0 = xyz[3 * p], x1 = xyz[3 * p + 1], x2 = xyz[3 * p + 2];
  double Xc[3], q[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)     // Mat44::transform: ((m0*x0 + m1*x1) + m2*x2) + m3, column-major storage
    Xc[i] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(Tc[i], x0), __dmul_rn(Tc[4 + i], x1)), __dmul_rn(Tc[8 + i], x2)), Tc[12 + i]);
#pragma unroll
  for (int i = 0; i < 3; ++i)
    q[i] = __dadd_rn(__dadd_rn(__dmul_rn(K[3 * i], Xc[0]), __dmul_rn(K[3 * i + 1], Xc[1])), __dmul_rn(K[3 * i + 2], Xc[2]));
  const double u = __ddiv_rn(q[0], q[2]), v = __ddiv_rn(q[1], q[2]);
  const int max_rows = rows - border - 1, max_cols = cols - border - 1;
  float sc = -2.0f;
  int r = -1, c = -1;
  if (isfinite(u) && isfinite(v) && fabs(u) < 1e9 && fabs(v) < 1e9) { r = (int)round(v); c = (int)round(u); }
  if (r >= border && r < max_rows && c >= border && c <= max_cols) {
    const float x = (float)u, y = (float)v;
    float d[25];
    int i = 0;
    for (int rr = -2; rr <= 2; ++rr)
      for (int cc = -2; cc <= 2; ++cc) d[i++] = interp2_u8(img, rows, cols, pitch, __fadd_rn((float)cc, x), __fadd_rn((float)rr, y));
    float sum = 0.f;
    for (int k = 0; k < 25; ++k) sum = __fadd_rn(sum, d[k]);
    const float mean = __fdiv_rn(sum, 25.0f);
    float ss = 0.f;
    for (int k = 0; k < 25; ++k) { d[k] = __fsub_rn(d[k], mean); ss = __fadd_rn(ss, __fmul_rn(d[k], d[k])); }
    const float norm = __fsqrt_rn(ss);
    const float den = __fmul_rn(ref_norm[p], norm);
    float dot = 0.f;
    for (int k = 0; k < 25; ++k) dot = __fadd_rn(dot, __fmul_rn(ref_patch[(size_t)p * 25 + k], d[k]));
    sc = ((double)den > 1e-6) ? __fdiv_rn(dot, den) : -1.0f;
  }
  score[p] = sc;
  rc[2 * p] = r; rc[2 * p + 1] = c;
}

// mask blocks around the re-observed points (src/photobundle.cc:533-536); mask: 1 = free
__global__ void k_mask_blocks(uint8_t* __restrict__ mask, int rows, int cols, int n, const int* __restrict__ rc, int radius) {
  const int side = 2 * radius + 1, per = side * side;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * per) return;
  const int p = (int)(t / per), e = (int)(t - (long long)p * per);
  const int r = rc[2 * p] + e / side - radius, c = rc[2 * p + 1] + e % side - radius;
  if (r >= 0 && r < rows && c >= 0 && c < cols) mask[(size_t)r * cols + c] = 0;
}

// new-point candidates (src/photobundle.cc:545-575): valid depth and IsLocalMax_ (src/imgproc.h:175-212) of the
// saliency map over the (2*nms+1)^2 neighbourhood with `>=`, masked pixels excluded; compacted with an atomic
// counter (the caller sorts them back into scan order)
__global__ void __launch_bounds__(256) k_candidates(const float* __restrict__ sal, const uint8_t* __restrict__ mask,
                                                    const float* __restrict__ depth, int rows, int cols, int border, int nms,
                                                    double min_depth, double max_depth, int capacity, int* __restrict__ count,
                                                    int* __restrict__ cand_rc, float* __restrict__ cand_sal) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int max_rows = rows - border - 1, max_cols = cols - border - 1;
  if (y < border || y >= max_rows || x < border || x >= max_cols) return;
  if (depth) {   // (no depth map on the device: the caller applies the depth test to what comes back)
    const float z = depth[(size_t)y * cols + x];
    if (!((double)z >= min_depth && (double)z <= max_depth)) return;
  }
  const float v = sal[(size_t)y * cols + x];
  if (nms > 0) {
    if (!mask[(size_t)y * cols + x] || v < 0.0f) return;
    for (int r = -nms; r <= nms; ++r)
      for (int c = -nms; c <= nms; ++c)
        if (!(!r && !c) && sal[(size_t)(y + r) * cols + x + c] >= v) return;
  }
  const int slot = atomicAdd(count, 1);
  if (slot < capacity) { cand_rc[2 * slot] = y; cand_rc[2 * slot + 1] = x; cand_sal[slot] = v; }
}

static dim3 grid_px(int rows, int cols) { return dim3((cols + 31) / 32, (rows + 7) / 8); }

cudaError_t launch_associate(const uint8_t* img, int rows, int cols, int pitch, int n, const double* xyz, const float* ref_patch,
                             const float* ref_norm, const double* Tc, const double* K, int border, float* score, int* rc,
                             cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  k_associate<<<(n + 127) / 128, 128, 0, stream>>>(img, rows, cols, pitch, n, xyz, ref_patch, ref_norm, Tc, K, border, score, rc);
  return cudaGetLastError();
}

cudaError_t launch_candidates(const float* sal, uint8_t* mask, const float* depth, int rows, int cols, int border, int nms,
                              int n_masked, const int* masked_rc, int mask_radius, double min_depth, double max_depth, int capacity,
                              int* count, int* cand_rc, float* cand_sal, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(mask, 1, (size_t)rows * cols, stream);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(count, 0, sizeof(int), stream);
  if (e != cudaSuccess) return e;
  if (n_masked > 0) {
    const long long tot = (long long)n_masked * (2 * mask_radius + 1) * (2 * mask_radius + 1);
    k_mask_blocks<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(mask, rows, cols, n_masked, masked_rc, mask_radius);
  }
  k_candidates<<<grid_px(rows, cols), 256, 0, stream>>>(sal, mask, depth, rows, cols, border, nms, min_depth, max_depth, capacity, count,
                                                         cand_rc, cand_sal);
  return cudaGetLastError();
}


cudaError_t launch_channels(int descriptor_type, const uint8_t* src, int rows, int cols, int spitch, uint8_t* scratch_a,
                            uint8_t* scratch_b, float* dst, int dpitch, size_t dplane, cudaStream_t stream) {
  const dim3 g = grid_px(rows, cols);
  if (descriptor_type == 1) {
    k_channels_ig<<<g, 256, 0, stream>>>(src, rows, cols, spitch, dst, dpitch, dplane);
  } else if (descriptor_type == 2) {
    k_gauss3_u8<<<g, 256, 0, stream>>>(src, rows, cols, spitch, scratch_a, spitch);
    k_census<<<g, 256, 0, stream>>>(scratch_a, rows, cols, spitch, scratch_b, spitch);
    k_bitplanes<<<g, 256, 0, stream>>>(scratch_b, rows, cols, spitch, dst, dpitch, dplane);
  } else {
    return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_saliency(const float* planes, int n_channels, int rows, int cols, int pitch, size_t plane, float* out,
                            cudaStream_t stream) {
  k_saliency<<<grid_px(rows, cols), 256, 0, stream>>>(planes, n_channels, rows, cols, pitch, plane, out);
  return cudaGetLastError();
}

cudaError_t launch_extract_patches(const float* planes, int n_channels, int rows, int cols, int pitch, size_t plane, int radius,
                                   int n, const int* xy, double* desc, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  k_extract_patches<<<n, 128, 0, stream>>>(planes, n_channels, rows, cols, pitch, plane, radius, n, xy, desc);
  return cudaGetLastError();
}

}  // namespace pba
