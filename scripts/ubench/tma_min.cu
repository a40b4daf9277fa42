// Minimal TMA 2-D box load (CUDA programming guide idiom, libcu++ wrappers): does a 16 B x 9 rows uint8 box work?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstdlib>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
template <bool GLOBAL_DESC>
__global__ void k(const __grid_constant__ CUtensorMap tmap_p, const CUtensorMap* tmap_g, int x, int y, unsigned* out, int nbytes) {
  const CUtensorMap* tm = GLOBAL_DESC ? tmap_g : &tmap_p;
  if (GLOBAL_DESC) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tm) : "memory");
  __shared__ alignas(128) unsigned char buf[16384];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(buf, tm, x, y, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, nbytes);
  } else {
    token = bar.arrive();
  }
  bar.wait(std::move(token));
  unsigned s = 0;
  for (int i = threadIdx.x; i < nbytes; i += blockDim.x) s += buf[i];
  atomicAdd(out, s);
}
int main(int argc, char** argv) {
  const int bw = argc > 1 ? atoi(argv[1]) : 16, bh = argc > 2 ? atoi(argv[2]) : 9;
  const int pitch = 1248, rows = 376 * 8;
  std::vector<uint8_t> h((size_t)pitch * rows);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 7 + (i >> 9));
  uint8_t* d; cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  unsigned* out; cudaMalloc(&out, 4); cudaMemset(out, 0, 4);
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  CUtensorMap tm;
  const cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)pitch};
  const cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh}, es[2] = {1, 1};
  CUresult r = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int)r);
  const int x = 101, y = 777;
  CUtensorMap* dtm; cudaMalloc(&dtm, sizeof(CUtensorMap)); cudaMemcpy(dtm, &tm, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
  const bool global_desc = argc > 3 && atoi(argv[3]);
  if (global_desc) k<true><<<1, 32>>>(tm, dtm, x, y, out, bw * bh); else k<false><<<1, 32>>>(tm, dtm, x, y, out, bw * bh);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned got; cudaMemcpy(&got, out, 4, cudaMemcpyDeviceToHost);
  unsigned want = 0; for (int rr = 0; rr < bh; ++rr) for (int c = 0; c < bw; ++c) want += h[(size_t)(y + rr) * pitch + x + c];
  printf("kernel: %s, sum %u, expected %u\n", cudaGetErrorString(e), got, want);
  return 0;
}
