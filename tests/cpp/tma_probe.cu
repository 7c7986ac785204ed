// TMA bring-up probe (TEST INFRASTRUCTURE): one cp.async.bulk.tensor.2d load of a [rows][cols] f32 tensor tile into shared memory, under
// several tensor-map settings. usage: tma_probe VARIANT   (one variant per process)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap m, float* out, int box_cols, int box_rows, int c0, int c1, int form) {
  extern __shared__ __align__(1024) unsigned char raw[];
  unsigned base = ((unsigned)__cvta_generic_to_shared(raw) + 1023u) & ~1023u;
  __shared__ __align__(8) unsigned long long bar;
  unsigned b = (unsigned)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(box_cols * box_rows * 4) : "memory");
    if (form == 0)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(base), "l"(&m), "r"(c0), "r"(c1), "r"(b) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(base), "l"(&m), "r"(c0), "r"(c1), "r"(b) : "memory");
  }
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(b) : "memory");
  const float* s = reinterpret_cast<const float*>(raw + (base - (unsigned)__cvta_generic_to_shared(raw)));
  for (int i = threadIdx.x; i < box_cols * box_rows; i += blockDim.x) out[i] = s[i];
}
int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { printf("no encode entry point\n"); return 1; }
  EncodeFn enc = (EncodeFn)p;
  const int rows = 200, cols = 1152;
  std::vector<float> h((size_t)rows * cols);
  for (size_t i = 0; i < h.size(); i++) h[i] = (float)i;
  float *d, *o; cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 128 * 32 * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  // variants: 0 = what conv_tc uses (box 32x128, SWIZZLE_128B, L2 128B, .tile form); 1 = same, form without .tile; 2 = no swizzle; 3 = box 32x8; 4 = no L2 promotion; 5 = coords (29, 128)
  int bc = 32, br = (variant == 3) ? 8 : 128, c0 = 0, c1 = 0, form = variant == 1 ? 1 : 0;
  CUtensorMapSwizzle sw = variant == 2 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B;
  CUtensorMapL2promotion l2 = variant == 4 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  if (variant == 5) { c0 = 29; c1 = 128; }
  alignas(64) CUtensorMap m;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)cols * 4};
  const cuuint32_t box[2] = {(cuuint32_t)bc, (cuuint32_t)br}, es[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("variant %d: encode rc=%d\n", variant, (int)r);
  if (r != CUDA_SUCCESS) return 1;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  k<<<1, 128, 40000>>>(m, o, bc, br, c0, c1, form);
  cudaError_t e = cudaDeviceSynchronize();
  printf("variant %d: kernel %s\n", variant, cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<float> g(bc * br); cudaMemcpy(g.data(), o, g.size() * 4, cudaMemcpyDeviceToHost);
  printf("  smem[0..7] = %g %g %g %g %g %g %g %g ; row1: %g %g ; row 9: %g\n", g[0], g[1], g[2], g[3], g[4], g[5], g[6], g[7], g[32], g[36], g[9 * 32]);
  return 0;
}
