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

static dim3 grid_px(int rows, int cols) { return dim3((cols + 31) / 32, (rows + 7) / 8); }

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
