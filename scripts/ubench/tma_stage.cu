// Footprint staging of K_A in isolation: Ampere-style cp.async (LDGSTS: 27 lanes x 4 bytes per footprint, what ships)
// against Hopper/Blackwell TMA (cp.async.bulk.tensor.2d: one elected lane, one 16 B x 9 rows box per footprint,
// completion through an mbarrier).  Same geometry as the product kernel: 14 warps per CTA, two CTAs per SM, one
// warp per "point", 8 footprints (one per observing frame) staged per batch from 8 L2-resident uint8 frames of
// 376 x 1248 bytes at random interior positions, then every lane reads its taps so that the data must have landed.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

constexpr int kWarps = 14, kSlots = 8, kRows = 9, kPitch = 1248, kImgRows = 376, kFrames = 8;

__device__ __forceinline__ void cp_async_4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}

template <int METHOD>   // 0: LDGSTS (12-byte rows, 108 B per footprint)   1: TMA (16-byte rows, 144 B, 256 B slots)
__global__ void __launch_bounds__(kWarps * 32, 2) k_stage(const uint8_t* frames, const __grid_constant__ CUtensorMap tmap,
                                                         const int2* origins /*[points][8] {row0 incl. frame, col0}*/, int n_points,
                                                         int iters, unsigned* sink, long long* cycles) {
  constexpr int kSlotBytes = METHOD == 0 ? 112 : 256;
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* fp = smem + warp * (kSlots * kSlotBytes);
  __shared__ __align__(8) unsigned long long bars[kWarps];
  if (METHOD >= 1 && lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bars[warp])));
  }
  __syncthreads();
  unsigned acc = 0, phase = 0;
  long long t_stage = 0;
  const int p0 = blockIdx.x * kWarps + warp;
  for (int it = 0; it < iters; ++it) {
    const int p = (p0 + it * gridDim.x * kWarps) % n_points;
    const int2 o = origins[p * 8 + (lane & 7)];
    const long long t0 = clock64();
    if (METHOD == 0) {
#pragma unroll
      for (int sl = 0; sl < kSlots; ++sl) {
        const int row0 = __shfl_sync(0xffffffffu, o.x, sl), cb = __shfl_sync(0xffffffffu, o.y, sl) & ~3;
        if (lane < 27) {
          const int r = lane / 3, w = lane - 3 * r;
          cp_async_4(fp + sl * kSlotBytes + 4 * lane, frames + (size_t)(row0 + r) * kPitch + cb + 4 * w);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
    } else if (METHOD == 2) {
      const unsigned bar = (unsigned)__cvta_generic_to_shared(&bars[warp]);
      if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kSlots * 288) : "memory");
      }
      __syncwarp();
#pragma unroll
      for (int sl = 0; sl < kSlots; ++sl) {
        const int row0 = __shfl_sync(0xffffffffu, o.x, sl), cb = __shfl_sync(0xffffffffu, o.y, sl) & ~15;
        if (lane < kRows) {
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 32, [%2];"
                       ::"r"((unsigned)__cvta_generic_to_shared(fp + sl * kSlotBytes + 32 * lane)), "l"(frames + (size_t)(row0 + lane) * kPitch + cb), "r"(bar) : "memory");
        }
      }
      unsigned done = 0;
      while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(phase) : "memory");
      }
      phase ^= 1;
    } else {
      const unsigned bar = (unsigned)__cvta_generic_to_shared(&bars[warp]);
      if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kSlots * 144) : "memory");
      }
#pragma unroll
      for (int sl = 0; sl < kSlots; ++sl) {
        const int row0 = __shfl_sync(0xffffffffu, o.x, sl), c0 = __shfl_sync(0xffffffffu, o.y, sl);
        if (lane == 0) {
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                       ::"r"((unsigned)__cvta_generic_to_shared(fp + sl * kSlotBytes)), "l"(&tmap), "r"(c0), "r"(row0), "r"(bar) : "memory");
        }
      }
      unsigned done = 0;
      while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(phase) : "memory");
      }
      phase ^= 1;
    }
    t_stage += clock64() - t0;
    // consume: every lane reads 12 taps of "its" pixel from two footprints (as the sampler does)
    constexpr int W = METHOD == 0 ? 12 : 16;
    const int py = (lane % 25) / 5, px = (lane % 25) % 5;
#pragma unroll
    for (int sl = 0; sl < kSlots; sl += 4) {
      const unsigned char* q = fp + sl * kSlotBytes + (py + 1) * W + px + 1 + (METHOD == 0 ? ((__shfl_sync(0xffffffffu, o.y, sl)) & 3) : METHOD == 2 ? ((__shfl_sync(0xffffffffu, o.y, sl)) & 15) : 0);
      acc += q[0] + q[1] + q[W] + q[W + 1] + q[-1] + q[2] + q[-W] + q[2 * W];
    }
    __syncwarp();
  }
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (lane == 0) atomicAdd((unsigned long long*)cycles, (unsigned long long)t_stage);
}

int main() {
  const int n_points = 4000, iters = 64;
  const size_t plane = (size_t)kImgRows * kPitch;
  std::vector<uint8_t> h(plane * kFrames);
  srand(7);
  for (auto& v : h) v = rand() & 255;
  std::vector<int2> org(n_points * 8);
  for (int p = 0; p < n_points; ++p) {
    const int r = 16 + rand() % (kImgRows - 48), c = 16 + rand() % (1241 - 48);
    for (int f = 0; f < 8; ++f) org[p * 8 + f] = make_int2(f * kImgRows + r + (rand() % 5) - 2, c + (rand() % 9) - 4);
  }
  uint8_t* d_fr; int2* d_org; unsigned* d_sink; long long* d_cyc;
  cudaMalloc(&d_fr, h.size() + 64); cudaMemcpy(d_fr, h.data(), h.size(), cudaMemcpyHostToDevice);
  cudaMalloc(&d_org, org.size() * sizeof(int2)); cudaMemcpy(d_org, org.data(), org.size() * sizeof(int2), cudaMemcpyHostToDevice);
  const int grid = 286;
  cudaMalloc(&d_sink, grid * kWarps * 32 * 4); cudaMalloc(&d_cyc, 8);
  // tensor map: uint8 [F * rows][pitch], box 16 x 9
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (!fn) { printf("cuTensorMapEncodeTiled not available\n"); return 1; }
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {(cuuint64_t)kPitch, (cuuint64_t)kImgRows * kFrames};
  const cuuint64_t strides[1] = {(cuuint64_t)kPitch};
  const cuuint32_t box[2] = {16, (cuuint32_t)kRows}, estr[2] = {1, 1};
  CUresult cr = ((EncodeFn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d_fr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)cr); return 1; }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mi = 0; mi < 3; ++mi) {
    const int method = mi == 0 ? 0 : mi == 1 ? 2 : 1;
    const size_t smem = (size_t)kWarps * kSlots * (method == 0 ? 112 : 256) + (method == 0 ? 90 * 1024 : 75 * 1024);   // pad to the product kernel's ~106 KB: two CTAs per SM
    if (method == 0) { cudaFuncSetAttribute(k_stage<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); cudaFuncSetAttribute(k_stage<0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); }
    else if (method == 1) { cudaFuncSetAttribute(k_stage<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); cudaFuncSetAttribute(k_stage<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); }
    else { cudaFuncSetAttribute(k_stage<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); cudaFuncSetAttribute(k_stage<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); }
    for (int rep = 0; rep < 3; ++rep) {
      cudaMemset(d_cyc, 0, 8);
      cudaEventRecord(e0);
      if (method == 0) k_stage<0><<<grid, kWarps * 32, smem>>>(d_fr, tmap, d_org, n_points, iters, d_sink, d_cyc);
      else if (method == 1) k_stage<1><<<grid, kWarps * 32, smem>>>(d_fr, tmap, d_org, n_points, iters, d_sink, d_cyc);
      else k_stage<2><<<grid, kWarps * 32, smem>>>(d_fr, tmap, d_org, n_points, iters, d_sink, d_cyc);
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      if (err != cudaSuccess) { printf("method %d (%s): %s\n", method, method == 1 ? "cp.async.bulk.tensor.2d" : method == 2 ? "cp.async.bulk" : "cp.async", cudaGetErrorString(err)); return 1; }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      long long cyc; cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
      if (rep == 2) {
        std::vector<unsigned> sk((size_t)grid * kWarps * 32);
        cudaMemcpy(sk.data(), d_sink, sk.size() * 4, cudaMemcpyDeviceToHost);
        unsigned long long chk = 0; for (unsigned v : sk) chk += v;
        printf("checksum of the staged taps %llu | ", chk);
      }
      if (rep == 2)
        printf("%s: %.1f us for %d batches per warp (%d warps) -> %.0f cycles per batch of 8 footprints (issue -> data visible), %.2f us per batch-wave; staging smem per warp %d B\n",
               method == 0 ? "LDGSTS (cp.async, 27 lanes x 4 B)      " : method == 2 ? "bulk copy (cp.async.bulk, 9 lanes x 32 B)" : "TMA (cp.async.bulk.tensor.2d, 16 B x 9)",
               ms * 1e3, iters, grid * kWarps, (double)cyc / ((double)grid * kWarps * iters), ms * 1e3 / iters, kSlots * (method == 0 ? 108 : 256));
    }
  }
  return 0;
}
