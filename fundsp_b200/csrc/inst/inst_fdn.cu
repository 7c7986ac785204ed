// Warp-per-voice reverb_stereo kernel (csrc/dsp/fdn_kernel.cuh) and its launcher.
#include "../dsp/fdn_kernel.cuh"
#include "../host/registry.h"
namespace fdsp { namespace host {
cudaError_t launch_fdn(const FdnArgs& a, int warps, cudaStream_t st) {
  const size_t smem = (size_t)warps * FDN_WARP_FLOATS * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(fdn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  const unsigned grid = (a.V + (unsigned)warps - 1) / (unsigned)warps;
  fdn_kernel<<<grid, 32 * warps, smem, st>>>(a);
  return cudaGetLastError();
}
int fdn_max_warps() { const int m = (int)((227 * 1024) / (FDN_WARP_FLOATS * sizeof(float))); return m < 10 ? m : 10; }   // __launch_bounds__(320)
}}
