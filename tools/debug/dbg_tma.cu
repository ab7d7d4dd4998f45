// standalone probe of the TMA patch load used by orb_describe_kernel
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -o dbg_tma dbg_tma.cu ; ./dbg_tma <variant>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
template <int RANK>
__device__ __forceinline__ void run(const CUtensorMap* tmap, int x0, int y0, int img, uint8_t* out, int bytes) {
  __shared__ __align__(128) uint8_t buf[4096];
  __shared__ __align__(8) unsigned long long bar;
  const uint32_t b = smem_u32(&bar), d = smem_u32(buf);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    if (RANK == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(d), "l"(tmap), "r"(b), "r"(x0), "r"(y0), "r"(img) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(d), "l"(tmap), "r"(b), "r"(x0), "r"(y0) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D%=;\nbra W%=;\nD%=:\n}" ::"r"(b), "r"(0) : "memory");
  for (int i = threadIdx.x; i < bytes; i += 32) out[i] = buf[i];
}
__global__ void k3(const __grid_constant__ CUtensorMap tmap, int x0, int y0, int img, uint8_t* out, int bytes) { run<3>(&tmap, x0, y0, img, out, bytes); }
__global__ void k2(const __grid_constant__ CUtensorMap tmap, int x0, int y0, int img, uint8_t* out, int bytes) { run<2>(&tmap, x0, y0, img, out, bytes); }
__global__ void k3g(const CUtensorMap* tmap, int x0, int y0, int img, uint8_t* out, int bytes) { run<3>(tmap, x0, y0, img, out, bytes); }
int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int pitch = 1280, rows = 376, imgs = 2;
  std::vector<uint8_t> h((size_t) pitch * rows * imgs);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t) (i * 7 + i / pitch);
  uint8_t *d, *o;
  cudaMalloc(&d, h.size());
  cudaMalloc(&o, 4096);
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  alignas(64) CUtensorMap tm;
  int bw = 32, bh = 31, rank = 3;
  CUtensorMapL2promotion l2 = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  if (variant == 1) rank = 2;
  if (variant == 2) l2 = CU_TENSOR_MAP_L2_PROMOTION_NONE;
  if (variant == 3) bh = 32;
  if (variant == 4) bw = 64;
  if (variant == 5) { bw = 64; bh = 32; rank = 2; l2 = CU_TENSOR_MAP_L2_PROMOTION_NONE; }
  const cuuint64_t dims3[3] = {pitch, rows, imgs}, strides3[2] = {pitch, (cuuint64_t) pitch * rows};
  const cuuint64_t dims2[2] = {pitch, (cuuint64_t) rows * imgs}, strides2[1] = {pitch};
  const cuuint32_t box[3] = {(cuuint32_t) bw, (cuuint32_t) bh, 1}, es[3] = {1, 1, 1};
  const CUresult r = ((EncodeFn) fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, rank == 3 ? dims3 : dims2, rank == 3 ? strides3 : strides2, box, es,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  const int x0 = (variant == 6) ? 96 : 101, y0 = 57, img = 1, bytes = bw * bh;
  if (variant == 8) {
    CUtensorMap* dtm;
    cudaMalloc(&dtm, sizeof(tm));
    cudaMemcpy(dtm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
    k3g<<<1, 32>>>(dtm, x0, y0, img, o, bytes);
  } else if (rank == 3) k3<<<1, 32>>>(tm, x0, y0, img, o, bytes);
  else k2<<<1, 32>>>(tm, x0, y0 + img * rows, 0, o, bytes);
  const cudaError_t e = cudaDeviceSynchronize();
  std::vector<uint8_t> res(bytes);
  cudaMemcpy(res.data(), o, bytes, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int y = 0; y < bh; ++y)
    for (int x = 0; x < bw; ++x) bad += res[y * bw + x] != h[((size_t) img * rows + y0 + y) * pitch + x0 + x];
  printf("variant %d: encode %d x0 %d sync: %s mismatches %d\n", variant, (int) r, x0, cudaGetErrorString(e), bad);
  return 0;
}
