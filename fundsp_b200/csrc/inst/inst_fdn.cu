// Warp-per-voice reverb_stereo kernel (csrc/dsp/fdn_kernel.cuh) and its launcher.
#include "../dsp/fdn_kernel_ts.cuh"
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
// time-split form: K warps per voice, `vpb` voices per CTA
template <int K> static cudaError_t launch_ts(const FdnArgs& a, int vpb, cudaStream_t st) {
  const size_t smem = (size_t)vpb * (2 * 32 * FDN2_RS + K * 2 * 32 * FDN2_PS + 2 * 2 * 64 + 2 * 64 + 3 * 32 + 32) * sizeof(float);
  static bool attr = false;
  if (!attr) { cudaError_t e = cudaFuncSetAttribute(fdn_kernel_ts<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); if (e != cudaSuccess) return e; attr = true; }
  const unsigned grid = (a.V + (unsigned)vpb - 1) / (unsigned)vpb;
  fdn_kernel_ts<K><<<grid, 32 * K * vpb, smem, st>>>(a, vpb);
  return cudaGetLastError();
}
cudaError_t launch_fdn_ts(const FdnArgs& a, int K, int vpb, cudaStream_t st) { return K == 4 ? launch_ts<4>(a, vpb, st) : launch_ts<2>(a, vpb, st); }
int fdn_ts_max_vpb(int K) {
  const size_t per = (size_t)(2 * 32 * FDN2_RS + K * 2 * 32 * FDN2_PS + 2 * 2 * 64 + 2 * 64 + 3 * 32 + 32) * sizeof(float);
  int m = (int)((227 * 1024) / per), t = 1024 / (32 * K);
  m = m < t ? m : t;
  return m < 8 ? m : 8;
}
int fdn_max_warps() { return (int)((227 * 1024) / (FDN_WARP_FLOATS * sizeof(float))) < 8 ? (int)((227 * 1024) / (FDN_WARP_FLOATS * sizeof(float))) : 8; }
}}
