// Plain-data arguments of the warp-per-voice reverb kernel (dsp/fdn_kernel.cuh), shared with the host runtime.
#pragma once
#ifndef __CUDACC_RTC__
#include <cstdint>
#endif

namespace fdsp {

struct FdnArgs {
  const uint32_t* params; uint32_t* state; const uint32_t* uniform;
  uint32_t p0, s0, u0;
  int scalar_row;            // >= 0: out = dry + P[scalar_row] * reverb  (Bus<MultiPass<2>, Unop<3, Reverb>>); -1: out = reverb
  const float* dry;          // stereo input rows: dry[v * dry_voice_stride + ch * dry_ch_stride + dry_offset + t]
  uint64_t dry_voice_stride; uint32_t dry_ch_stride, dry_offset;
  float* out; const uint32_t* row_map; uint32_t out_stride, out_offset;  // per-voice rows (2 per voice) or null
  float* partial;            // [grid][2][n] CTA partial mixes or null
  float* ring; uint64_t ring_voice_stride;
  uint32_t V, n;
};

}  // namespace fdsp
