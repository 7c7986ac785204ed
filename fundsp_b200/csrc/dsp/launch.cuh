// Launcher template + registry entry macro shared by the ahead-of-time instance files (csrc/inst/*.cu).
#pragma once
#include "../host/registry.h"
#include "bank_kernel.cuh"
#include "bank_kernel_ws.cuh"

namespace fdsp {
namespace host {

constexpr int NT = 128;

template <class G, int MODE, bool TB> cudaError_t launch_one(const BankArgs& a, unsigned grid, size_t smem, cudaStream_t st) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(bank_kernel<G, NT, MODE, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  bank_kernel<G, NT, MODE, TB><<<grid, NT, smem, st>>>(a);
  return cudaGetLastError();
}
template <class G, bool TB> cudaError_t launch_mode(const BankArgs& a, int mode, size_t table_bytes, cudaStream_t st) {
  const unsigned vpc = a.vpc ? a.vpc : (unsigned)NT, grid = (a.V + vpc - 1) / vpc;
  const size_t smem = ((mode & 2) ? sizeof(float) * mix_tile_floats(G::OUT, NT) : 0) + (TB ? table_bytes : 0);
  switch (mode & 3) {
    case 1: return launch_one<G, 1, TB>(a, grid, smem, st);
    case 2: return launch_one<G, 2, TB>(a, grid, smem, st);
    case 3: return launch_one<G, 3, TB>(a, grid, smem, st);
    default: return cudaErrorInvalidValue;
  }
}
// warp-specialised variant (bank_kernel_ws.cuh): 2*NT threads per CTA, stage A and stage B of the top-level Pipe in different warps
template <class G, int MODE, bool TB> cudaError_t launch_ws_one(const BankArgs& a, unsigned grid, size_t smem, cudaStream_t st) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(bank_kernel_ws<G, NT, MODE, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  bank_kernel_ws<G, NT, MODE, TB><<<grid, 2 * NT, smem, st>>>(a);
  return cudaGetLastError();
}
template <class G, bool TB> cudaError_t launch_ws_mode(const BankArgs& a, int mode, size_t smem, cudaStream_t st) {
  const unsigned vpc = a.vpc ? a.vpc : (unsigned)NT, grid = (a.V + vpc - 1) / vpc;
  switch (mode & 3) {
    case 1: return launch_ws_one<G, 1, TB>(a, grid, smem, st);
    case 2: return launch_ws_one<G, 2, TB>(a, grid, smem, st);
    case 3: return launch_ws_one<G, 3, TB>(a, grid, smem, st);
    default: return cudaErrorInvalidValue;
  }
}
// table_bytes > 0 asks for the shared-memory wavetable variant (only meaningful when G reads a wavetable and it fits).
// mode bit 2 (value 4) asks for the warp-specialised kernel where the graph has one and its shared memory fits.
template <class G> cudaError_t launch_t(const BankArgs& a, int mode, size_t table_bytes, cudaStream_t st) {
  if constexpr (WsOk<G>::value) {
    if (mode & 4) {
      typedef typename PipeParts<G>::A A;
      const bool mix = (mode & 2) != 0;
      const size_t base = (mix ? sizeof(float) * mix_tile_floats(G::OUT, NT) : 0) + sizeof(float) * ws_hand_floats(A::OUT, G::OUT, NT, mix);
      const bool want_tb = WaveKind<G>::value >= 0 && table_bytes > 0;
      if (want_tb && base + table_bytes <= 227 * 1024) return launch_ws_mode<G, (WaveKind<G>::value >= 0)>(a, mode, base + table_bytes, st);
      if (!want_tb && base <= 227 * 1024) return launch_ws_mode<G, false>(a, mode, base, st);
    }
  }
  if (WaveKind<G>::value >= 0 && table_bytes > 0) {
    const size_t smem = ((mode & 2) ? sizeof(float) * mix_tile_floats(G::OUT, NT) : 0) + table_bytes;
    if (smem <= 227 * 1024) return launch_mode<G, (WaveKind<G>::value >= 0)>(a, mode, table_bytes, st);
  }
  return launch_mode<G, false>(a, mode, 0, st);
}
template <class G> int wave_kind_t() { return WaveKind<G>::value; }
inline int threads_t() { return NT; }

#define FDSP_REG(...) \
  {#__VA_ARGS__, __VA_ARGS__::IN, __VA_ARGS__::OUT, __VA_ARGS__::NP, __VA_ARGS__::NS, __VA_ARGS__::NU, &launch_t<__VA_ARGS__>, &threads_t, &wave_kind_t<__VA_ARGS__>}
#define FDSP_INSTANCES(name, ...)                     \
  extern const KernelEntry kInst_##name[] = {__VA_ARGS__}; \
  extern const int kInst_##name##_n = (int)(sizeof(kInst_##name) / sizeof(kInst_##name[0]));

// ---- shorthand for the config graphs (expanded to canonical type expressions by registry.cpp)
typedef Pipe<Constant<1>, Sine> SineHz;
typedef Pipe<Constant<1>, WaveSynth<0, 1>> SawHz;
typedef Pipe<Unop<1, Unop<3, Unop<3, SineHz>>>, Sine> Fm;                       // sine_hz(f)*f*m+f >> sine()
typedef Pipe<Pipe<MultiSplit<2, 16>, Feedback<1, Multi<30, 0, 32, Pipe<Delay, Fir<3>>>>>,
             Binop<2, Multi<31, 0, 32, Panner<1>>, Constant<2>>> ReverbStereo;  // src/prelude.rs:1732-1762
typedef Pipe<Binop<2, Pipe<Stack<SawHz, Constant<2>>, Moog<3>>, AdsrLive>, Panner<1>> SubtractiveDry;
typedef Pipe<SubtractiveDry, Bus<MultiPass<2>, Unop<3, ReverbStereo>>> SubtractiveVoice;  // config 4

}  // namespace host
}  // namespace fdsp
