// fundsp_b200 warp-per-voice kernel for `reverb_stereo` (reference src/prelude.rs:1732-1762): the 32-line Hadamard
// feedback delay network  multisplit<2,16> >> fdn<32>(stacki<32>(delay >> fir3)) >> sumf<32>(pan) * dc((1/16,1/16)).
//
// The reference ticks the whole inner graph per sample (src/feedback.rs:136-146): 32 x (Delay::tick src/delay.rs:116-124
// + Fir<U3>::tick src/fir.rs:57-70) + one 32-point Hadamard (src/feedback.rs:35-57). Here ONE WARP evaluates one voice:
//   lane l   = delay line l (its FIR shift register, feedback value, pan weights, ring index live in registers)
//   Hadamard = 5 butterfly stages of __shfl_xor (upper lane computes partner - own, exactly the reference's x - y)
//   delay lines live in HBM, voice-major [voice][line][ring]; per 64-sample block the warp stages the 32 x 64 ring
//   slice it will read into shared memory with coalesced cp.async (double buffered, next block prefetched while the
//   current one is computed: the shortest line is >= 130 samples) and writes the 32 x 64 new samples back with
//   coalesced stores. This is the one HBM-bound program of the path: 32 lines x (4 B read + 4 B write) per voice-sample.
//   The output `Reduce` (sum of 32 pans, left fold in index order, src/audionode.rs:2442-2463) is done from a shared
//   transpose so the sum order — and therefore every bit — matches the reference.
// The wet/dry composition around it,  dry >> (multipass::<U2>() & s * reverb_stereo(..)), is applied in the epilogue.
// Word layout of the reverb inside the class arrays (DFS order, see csrc/host/graph.cpp):
//   P: 32 x Fir weights(3) | 32 x Panner(lw, rw) | Constant<2>          (162 words, first row p0)
//   S: Feedback value[32] | 32 x (Delay idx, Fir v[3])                   (160 words, first row s0)
//   U: 32 x ring length                                                  (first row u0)
#pragma once
#include "math.cuh"
#include "fdn_args.h"

namespace fdsp {



constexpr int FDN_RS = 65;                 // padded row stride (floats): bank = (line + t) % 32, conflict-free both ways
constexpr int FDN_PS = 33;
constexpr int FDN_WARP_FLOATS = 2 * 32 * FDN_RS + 2 * 32 * FDN_PS + 2 * 2 * 64 + 2 * 64 + 3 * 32;  // rbuf x2, pbuf, dbuf x2, obuf, line tables

FDSP_DEV void cp_async4(uint32_t smem, const float* g) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem), "l"(g) : "memory"); }
FDSP_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> FDSP_DEV void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// blockDim.x = 32 * W (W voices per CTA), dynamic smem = W * FDN_WARP_FLOATS * 4 bytes
__global__ void __launch_bounds__(256) fdn_kernel(const FdnArgs a) {
  extern __shared__ __align__(16) float fdn_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  const uint32_t v = blockIdx.x * W + warp;
  const bool active = v < a.V;
  float* sm = fdn_smem + (size_t)warp * FDN_WARP_FLOATS;
  float* rbuf = sm;                                  // [2][32][65] ring slices read this / next block; reused for the new samples
  float* pbuf = rbuf + 2 * 32 * FDN_RS;              // [2][32][33] pan products of a 32-sample half block
  float* dbuf = pbuf + 2 * 32 * FDN_PS;              // [2][2][64] stereo input of this / next block
  float* obuf = dbuf + 2 * 2 * 64;                   // [2][64] final output of this block (for the CTA mix)
  uint32_t* tlen = reinterpret_cast<uint32_t*>(obuf + 2 * 64);  // [32] ring length, [32] ring offset, [32] ring index
  uint32_t* toff = tlen + 32;
  uint32_t* tidx = toff + 32;

  float w0 = 0, w1 = 0, w2 = 0, lw = 0, rw = 0, c0 = 0, c1 = 0, scalar = 1.0f, value = 0, f0 = 0, f1 = 0, f2 = 0;
  uint32_t idx = 0, len = 1, off = 0;
  float* ring = nullptr;
  if (active) {
    const uint32_t V = a.V;
    auto P = [&](uint32_t row) { return __uint_as_float(__ldg(a.params + (size_t)row * V + v)); };
    auto S = [&](uint32_t row) { return a.state[(size_t)row * V + v]; };
    w0 = P(a.p0 + 3 * lane); w1 = P(a.p0 + 3 * lane + 1); w2 = P(a.p0 + 3 * lane + 2);
    lw = P(a.p0 + 96 + 2 * lane); rw = P(a.p0 + 96 + 2 * lane + 1);
    c0 = P(a.p0 + 160); c1 = P(a.p0 + 161);
    if (a.scalar_row >= 0) scalar = P((uint32_t)a.scalar_row);
    value = __uint_as_float(S(a.s0 + lane));
    idx = S(a.s0 + 32 + 4 * lane);
    f0 = __uint_as_float(S(a.s0 + 32 + 4 * lane + 1)); f1 = __uint_as_float(S(a.s0 + 32 + 4 * lane + 2)); f2 = __uint_as_float(S(a.s0 + 32 + 4 * lane + 3));
    len = __ldg(a.uniform + a.u0 + lane);
    uint32_t incl = len;  // inclusive scan over lanes -> ring offset of each line
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
    off = incl - len;
    ring = a.ring + (size_t)v * a.ring_voice_stride;
    tlen[lane] = len; toff[lane] = off; tidx[lane] = idx;
  }
  __syncwarp();
  const uint32_t rb0 = (uint32_t)__cvta_generic_to_shared(rbuf), db0 = (uint32_t)__cvta_generic_to_shared(dbuf);
  const float* dry = active ? a.dry + (size_t)v * a.dry_voice_stride + a.dry_offset : nullptr;

  // stage the ring slice + stereo input a block will read: lanes sweep time, lines are looped (coalesced 128 B rows)
  auto prefetch = [&](int buf, uint32_t t0, int nb, uint32_t adv) {
    if (active && nb > 0) {
#pragma unroll 4
      for (int l = 0; l < 32; l++) {
        const uint32_t L = tlen[l], base = toff[l], i0 = tidx[l] + adv;   // ring index at the start of that block
#pragma unroll
        for (int q = 0; q < 2; q++) {
          const int t = lane + 32 * q;
          if (t < nb) {
            uint32_t pos = i0 + 1u + (uint32_t)t;           // Delay::tick reads buffer[i + 1] after writing buffer[i]
            pos -= (pos >= L) ? L : 0u; pos -= (pos >= L) ? L : 0u;
            cp_async4(rb0 + 4u * (uint32_t)((buf * 32 + l) * FDN_RS + t), ring + base + pos);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const int t = lane + 32 * q;
        if (t < nb) {
          cp_async4(db0 + 4u * (uint32_t)((buf * 2 + 0) * 64 + t), dry + t0 + t);
          cp_async4(db0 + 4u * (uint32_t)((buf * 2 + 1) * 64 + t), dry + a.dry_ch_stride + t0 + t);
        }
      }
    }
    cp_async_commit();
  };

  const float hz = (float)(1.0 / sqrt(32.0));
  uint32_t sgn[5];
#pragma unroll
  for (int s = 0; s < 5; s++) sgn[s] = (lane & (1 << s)) ? 0x80000000u : 0u;
  int cur = 0;
  prefetch(0, 0u, (int)(a.n < 64u ? a.n : 64u), 0u);
#pragma unroll 1
  for (uint32_t t0 = 0; t0 < a.n; t0 += 64) {
    const int nb = (a.n - t0) < 64u ? (int)(a.n - t0) : 64;
    const uint32_t rest = a.n - t0 - (uint32_t)nb;
    prefetch(cur ^ 1, t0 + (uint32_t)nb, (int)(rest < 64u ? rest : 64u), (uint32_t)nb);
    cp_async_wait<1>();
    __syncwarp();
    if (active) {
      float* rb = rbuf + (cur * 32 + lane) * FDN_RS;
      const float* din = dbuf + (cur * 2 + (lane & 1)) * 64;   // MultiSplit<2,16>: channel c reads input c % 2
#pragma unroll 1
      for (int h0 = 0; h0 < nb; h0 += 32) {
        const int hn = (nb - h0) < 32 ? (nb - h0) : 32;
        int tt = 0;
        // groups of 8 samples: all shared-memory reads first, then 8 interleaved Hadamard chains, then the stores
        // (consecutive samples are independent: the feedback value only reaches the ring after >= 129 samples)
#pragma unroll 1
        for (; tt + 8 <= hn; tt += 8) {
          float d[8], x[8], o[8], h[8];
#pragma unroll
          for (int u = 0; u < 8; u++) { d[u] = rb[h0 + tt + u]; x[u] = din[h0 + tt + u]; }
#pragma unroll
          for (int u = 0; u < 8; u++) { f0 = f1; f1 = f2; f2 = d[u]; o[u] = (w0 * f0 + w1 * f1) + w2 * f2; h[u] = o[u]; }
#pragma unroll
          for (int s = 0; s < 5; s++) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
              const float y = __shfl_xor_sync(0xffffffffu, h[u], 1 << s);
              h[u] = y + __uint_as_float(__float_as_uint(h[u]) ^ sgn[s]);   // upper lane: partner - own; lower lane: own + partner
            }
          }
#pragma unroll
          for (int u = 0; u < 8; u++) {
            rb[h0 + tt + u] = x[u] + value;    // Feedback: x.tick(input + value); Delay stores it at ring[i]
            value = h[u] * hz;
            pbuf[(0 * 32 + lane) * FDN_PS + tt + u] = o[u] * lw;   // Panner<U1>::process: input * weight
            pbuf[(1 * 32 + lane) * FDN_PS + tt + u] = o[u] * rw;
          }
        }
#pragma unroll 1
        for (; tt < hn; tt++) {
          const int t = h0 + tt;
          const float d = rb[t];                 // Delay output for this sample
          rb[t] = din[t] + value;
          f0 = f1; f1 = f2; f2 = d;              // Fir<U3> shift register
          const float o = (w0 * f0 + w1 * f1) + w2 * f2;
          float h = o;                           // FrameHadamard<U32>
#pragma unroll
          for (int s = 0; s < 5; s++) {
            const float y = __shfl_xor_sync(0xffffffffu, h, 1 << s);
            h = y + __uint_as_float(__float_as_uint(h) ^ sgn[s]);
          }
          value = h * hz;
          pbuf[(0 * 32 + lane) * FDN_PS + tt] = o * lw;
          pbuf[(1 * 32 + lane) * FDN_PS + tt] = o * rw;
        }
        __syncwarp();
        if (lane < hn) {                         // Reduce<U32, Panner, FrameAdd>: left fold in channel order, then * dc, * s, + dry
          const int t = h0 + lane;
          float sl = pbuf[(0 * 32 + 0) * FDN_PS + lane], sr = pbuf[(1 * 32 + 0) * FDN_PS + lane];
#pragma unroll
          for (int l = 1; l < 32; l++) { sl += pbuf[(0 * 32 + l) * FDN_PS + lane]; sr += pbuf[(1 * 32 + l) * FDN_PS + lane]; }
          sl = sl * c0; sr = sr * c1;
          if (a.scalar_row >= 0) {
            sl = dbuf[(cur * 2 + 0) * 64 + t] + sl * scalar;
            sr = dbuf[(cur * 2 + 1) * 64 + t] + sr * scalar;
          }
          obuf[t] = sl; obuf[64 + t] = sr;
          if (a.out) {
            float* orow = a.out + (size_t)__ldg(a.row_map + v) * a.out_stride + a.out_offset + t0 + t;
            orow[0] = sl; orow[a.out_stride] = sr;
          }
        }
        __syncwarp();
      }
      // write the 32 x nb new samples back to the rings (coalesced rows) and advance the ring indices
#pragma unroll 4
      for (int l = 0; l < 32; l++) {
        const uint32_t L = tlen[l], base = toff[l], i0 = tidx[l];
#pragma unroll
        for (int q = 0; q < 2; q++) {
          const int t = lane + 32 * q;
          if (t < nb) {
            uint32_t pos = i0 + (uint32_t)t;
            pos -= (pos >= L) ? L : 0u;
            ring[base + pos] = rbuf[(cur * 32 + l) * FDN_RS + t];
          }
        }
      }
      __syncwarp();
      idx += (uint32_t)nb; idx -= (idx >= len) ? len : 0u;
      tidx[lane] = idx;
    } else {
      if (a.partial) { for (int t = lane; t < 128; t += 32) obuf[t] = 0.0f; }
    }
    if (a.partial) {   // CTA mix in warp (= voice) order, deterministic
      __syncthreads();
      for (int e = threadIdx.x; e < 2 * nb; e += blockDim.x) {
        const int ch = e / nb, t = e - ch * nb;
        float s = fdn_smem[(size_t)0 * FDN_WARP_FLOATS + (obuf - sm) + ch * 64 + t];
        for (int w = 1; w < W; w++) s += fdn_smem[(size_t)w * FDN_WARP_FLOATS + (obuf - sm) + ch * 64 + t];
        a.partial[((size_t)blockIdx.x * 2 + ch) * a.n + t0 + t] = s;
      }
      __syncthreads();
    }
    __syncwarp();
    cur ^= 1;
  }
  cp_async_wait<0>();
  if (active) {
    const uint32_t V = a.V;
    a.state[(size_t)(a.s0 + lane) * V + v] = __float_as_uint(value);
    a.state[(size_t)(a.s0 + 32 + 4 * lane) * V + v] = idx;
    a.state[(size_t)(a.s0 + 32 + 4 * lane + 1) * V + v] = __float_as_uint(f0);
    a.state[(size_t)(a.s0 + 32 + 4 * lane + 2) * V + v] = __float_as_uint(f1);
    a.state[(size_t)(a.s0 + 32 + 4 * lane + 3) * V + v] = __float_as_uint(f2);
  }
}

}  // namespace fdsp
