// Plain-data arguments of the warp-per-voice reverb kernel (dsp/fdn_kernel.cuh), shared with the host runtime.
#pragma once
#ifndef __CUDACC_RTC__
#include <cstdint>
#endif

namespace fdsp {

// physical ring length of a delay line of `len` samples: a multiple of 64 floats, so block-aligned slices are 16-byte aligned and a
// block never straddles the end (the kernel and the host's allocation both use this)
#ifdef __CUDACC__
__host__ __device__
#endif
constexpr uint32_t fdn_ring_phys(uint32_t len) { return (len + 63u) & ~63u; }
constexpr uint32_t FDN_MIN_RING = 193;   // shortest delay line the kernel takes (prefetch distance 2 blocks + the block itself)

struct FdnArgs {
  const uint32_t* params; uint32_t* state; const uint32_t* uniform;
  uint32_t p0, s0, u0;
  int scalar_row;            // >= 0: out = dry + P[scalar_row] * reverb  (Bus<MultiPass<2>, Unop<3, Reverb>>); -1: out = reverb
  const float* dry;          // stereo input rows: dry[v * dry_voice_stride + ch * dry_ch_stride + dry_offset + t]
  uint64_t dry_voice_stride; uint32_t dry_ch_stride, dry_offset;
  float* out; const uint32_t* row_map; uint32_t out_stride, out_offset;  // per-voice rows (2 per voice) or null
  float* partial;            // [V][2][n] per-voice rows of the mix-down (mix_reduce_kernel adds them in voice order) or null
  float* ring; uint64_t ring_voice_stride;
  uint32_t V, n;
  uint32_t flags;            // diagnostics (FDSP_FDN_FLAGS): bit 0 = one sample per lane and pass, bit 1 = three-stage form
};

}  // namespace fdsp
