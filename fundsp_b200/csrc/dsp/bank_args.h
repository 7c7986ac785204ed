// Plain-data kernel arguments shared by host (csrc/host) and device (csrc/dsp) code.
#pragma once
#ifndef __CUDACC_RTC__
#include <cstdint>
#endif

namespace fdsp {

// Wavetable set for one waveform (reference src/wavetable.rs:82-123): up to 48 tables, ascending pitch.
struct WaveTableDev {
  int n;
  float pitch[48];
  int off[48];
  int len[48];
  int total;          // floats in `data`
  const float* data;
};

struct BankArgs {
  const uint32_t* params;   // [NP][V]
  uint32_t* state;          // [NS][V]
  const uint32_t* uniform;  // [NU]
  float* dline;             // [sum(len)][V]
  const WaveTableDev* wt;   // [6]
  const float* in;          // shared bank input [IN][in_stride] or null
  float* out;               // per-voice output rows or null
  float* partial;           // [grid][OUT][n] per-CTA mix partials or null
  uint32_t V, n;
  uint32_t in_stride, in_offset;
  uint32_t out_stride, out_offset;  // row stride / first sample (multiples of 4 for the vector path)
  const uint32_t* row_map;          // class-local voice -> first output row (voice-major rows) inside `out`
  float sr, sd64, sd32;
};

}  // namespace fdsp
