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

}  // namespace pba
