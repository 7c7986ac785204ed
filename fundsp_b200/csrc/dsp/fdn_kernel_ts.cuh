// fundsp_b200 `reverb_stereo` kernel, time-split form: K warps per voice (see fdn_kernel.cuh for the algorithm and the
// reference citations). Inside a 64-sample block the feedback only reaches a ring after >= 129 samples, so the samples of
// a block are independent given the ring slice: warp k of a voice evaluates the sample range [nb*k/K, nb*(k+1)/K) for all
// 32 lines (lane = line). What a warp needs from outside its range is read-only history: the 3 previous delay outputs
// (FIR shift register + the Hadamard of the sample just before its range, recomputed redundantly to get the feedback
// value that is added to its first input). K x more warps per voice hide the shuffle / shared-memory latency that bounds
// the single-warp form when only ~1000 voices exist; prefetch and write-back are split across the K warps by line.
#pragma once
#include "fdn_kernel.cuh"

namespace fdsp {

constexpr int FDN2_RS = 69;   // 3 history columns + 64 + pad: (line * 69 + t) % 32 = (5 * line + t) % 32, conflict-free over lines
constexpr int FDN2_PS = 9;    // pan products are reduced every 8 samples
FDSP_DEV int fdn2_voice_floats(int K) { return 2 * 32 * FDN2_RS + K * 2 * 32 * FDN2_PS + 2 * 2 * 64 + 2 * 64 + 3 * 32 + 32; }

FDSP_DEV void named_bar(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// blockDim.x = 32 * K * VPB (VPB voices per CTA); dynamic smem = VPB * fdn2_voice_floats(K) * 4
template <int K>
__global__ void __launch_bounds__(32 * K * 8) fdn_kernel_ts(const FdnArgs a, int VPB) {
  extern __shared__ __align__(16) float fdn_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int vin = warp / K, k = warp - vin * K;          // voice slot in the CTA, time slice of this warp
  const uint32_t v = blockIdx.x * VPB + vin;
  const bool active = v < a.V;
  const int VF = fdn2_voice_floats(K);
  float* sm = fdn_smem + (size_t)vin * VF;
  float* rbuf = sm;                                      // [2][32][69]: columns 0..2 history (d[-3..-1]), 3.. the slice
  float* pbuf = rbuf + 2 * 32 * FDN2_RS + k * 2 * 32 * FDN2_PS;   // this warp's [2][32][9]
  float* dbuf = rbuf + 2 * 32 * FDN2_RS + K * 2 * 32 * FDN2_PS;   // [2][2][64]
  float* obuf = dbuf + 2 * 2 * 64;                       // [2][64]
  uint32_t* tlen = reinterpret_cast<uint32_t*>(obuf + 2 * 64);
  uint32_t* toff = tlen + 32;
  uint32_t* tidx = toff + 32;
  float* vbuf = reinterpret_cast<float*>(tidx + 32);     // [32] feedback value after the last sample of the previous block
  const int bar_id = 1 + vin, bar_n = 32 * K;

  float w0 = 0, w1 = 0, w2 = 0, lw = 0, rw = 0, c0 = 0, c1 = 0, scalar = 1.0f;
  float* ring = nullptr;
  if (active) {
    const uint32_t V = a.V;
    auto P = [&](uint32_t row) { return __uint_as_float(__ldg(a.params + (size_t)row * V + v)); };
    auto S = [&](uint32_t row) { return a.state[(size_t)row * V + v]; };
    w0 = P(a.p0 + 3 * lane); w1 = P(a.p0 + 3 * lane + 1); w2 = P(a.p0 + 3 * lane + 2);
    lw = P(a.p0 + 96 + 2 * lane); rw = P(a.p0 + 96 + 2 * lane + 1);
    c0 = P(a.p0 + 160); c1 = P(a.p0 + 161);
    if (a.scalar_row >= 0) scalar = P((uint32_t)a.scalar_row);
    ring = a.ring + (size_t)v * a.ring_voice_stride;
    if (k == 0) {
      const uint32_t len = __ldg(a.uniform + a.u0 + lane);
      uint32_t incl = len;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
      tlen[lane] = len; toff[lane] = incl - len; tidx[lane] = S(a.s0 + 32 + 4 * lane);
      vbuf[lane] = __uint_as_float(S(a.s0 + lane));
      // FIR history d[-3], d[-2], d[-1] of the first block = the saved shift register (v[0], v[1], v[2])
      float* h = rbuf + (0 * 32 + lane) * FDN2_RS;
      h[0] = __uint_as_float(S(a.s0 + 32 + 4 * lane + 1)); h[1] = __uint_as_float(S(a.s0 + 32 + 4 * lane + 2)); h[2] = __uint_as_float(S(a.s0 + 32 + 4 * lane + 3));
    }
  }
  named_bar(bar_id, bar_n);
  const uint32_t rb0 = (uint32_t)__cvta_generic_to_shared(rbuf), db0 = (uint32_t)__cvta_generic_to_shared(dbuf);
  const float* dry = active ? a.dry + (size_t)v * a.dry_voice_stride + a.dry_offset : nullptr;
  const int l0 = (32 * k) / K, l1 = (32 * (k + 1)) / K;   // lines this warp stages / writes back

  auto prefetch = [&](int buf, uint32_t t0, int nb, uint32_t adv) {
    if (active && nb > 0) {
#pragma unroll 4
      for (int l = l0; l < l1; l++) {
        const uint32_t L = tlen[l], base = toff[l], i0 = tidx[l] + adv;
#pragma unroll
        for (int q = 0; q < 2; q++) {
          const int t = lane + 32 * q;
          if (t < nb) {
            uint32_t pos = i0 + 1u + (uint32_t)t;
            pos -= (pos >= L) ? L : 0u; pos -= (pos >= L) ? L : 0u;
            cp_async4(rb0 + 4u * (uint32_t)((buf * 32 + l) * FDN2_RS + 3 + t), ring + base + pos);
          }
        }
      }
      if (k == 0) {
#pragma unroll
        for (int q = 0; q < 2; q++) {
          const int t = lane + 32 * q;
          if (t < nb) {
            cp_async4(db0 + 4u * (uint32_t)((buf * 2 + 0) * 64 + t), dry + t0 + t);
            cp_async4(db0 + 4u * (uint32_t)((buf * 2 + 1) * 64 + t), dry + a.dry_ch_stride + t0 + t);
          }
        }
      }
    }
    cp_async_commit();
  };

  const float hz = (float)(1.0 / sqrt(32.0));
  uint32_t sgn[5];
#pragma unroll
  for (int s = 0; s < 5; s++) sgn[s] = (lane & (1 << s)) ? 0x80000000u : 0u;
  auto hadamard1 = [&](float h) {
#pragma unroll
    for (int s = 0; s < 5; s++) { const float y = __shfl_xor_sync(0xffffffffu, h, 1 << s); h = y + __uint_as_float(__float_as_uint(h) ^ sgn[s]); }
    return h;
  };

  int cur = 0;
  prefetch(0, 0u, (int)(a.n < 64u ? a.n : 64u), 0u);
#pragma unroll 1
  for (uint32_t t0 = 0; t0 < a.n; t0 += 64) {
    const int nb = (a.n - t0) < 64u ? (int)(a.n - t0) : 64;
    const uint32_t rest = a.n - t0 - (uint32_t)nb;
    prefetch(cur ^ 1, t0 + (uint32_t)nb, (int)(rest < 64u ? rest : 64u), (uint32_t)nb);
    cp_async_wait<1>();
    named_bar(bar_id, bar_n);                             // the whole slice (all lines, stereo input) is in shared memory
    const int ta = (nb * k) / K, tb = (nb * (k + 1)) / K; // this warp's sample range
    float* rb = rbuf + (cur * 32 + lane) * FDN2_RS + 3;   // rb[t] = delay output d[t], rb[-3..-1] = history
    const float* din = dbuf + (cur * 2 + (lane & 1)) * 64;
    float f0 = 0, f1 = 0, f2 = 0, value = 0;
    if (active && tb > ta) {
      // history before the range: FIR registers and the feedback value entering sample ta
      const float dm3 = rb[ta - 3];
      f0 = dm3; f1 = rb[ta - 2]; f2 = rb[ta - 1];
      if (ta == 0) value = vbuf[lane];
      else value = hadamard1((w0 * dm3 + w1 * f1) + w2 * f2) * hz;   // value after sample ta-1, recomputed
    }
    named_bar(bar_id, bar_n);                             // everybody holds its history: the slice may now be overwritten in place
    if (active && tb > ta) {
      float* orow = a.out ? a.out + (size_t)__ldg(a.row_map + v) * a.out_stride + a.out_offset + t0 : nullptr;
#pragma unroll 1
      for (int g0 = ta; g0 < tb; g0 += 8) {
        const int gn = (tb - g0) < 8 ? (tb - g0) : 8;
        if (gn == 8) {
          float d[8], x[8], o[8], h[8];
#pragma unroll
          for (int u = 0; u < 8; u++) { d[u] = rb[g0 + u]; x[u] = din[g0 + u]; }
#pragma unroll
          for (int u = 0; u < 8; u++) { f0 = f1; f1 = f2; f2 = d[u]; o[u] = (w0 * f0 + w1 * f1) + w2 * f2; h[u] = o[u]; }
#pragma unroll
          for (int s = 0; s < 5; s++) {
#pragma unroll
            for (int u = 0; u < 8; u++) { const float y = __shfl_xor_sync(0xffffffffu, h[u], 1 << s); h[u] = y + __uint_as_float(__float_as_uint(h[u]) ^ sgn[s]); }
          }
#pragma unroll
          for (int u = 0; u < 8; u++) {
            rb[g0 + u] = x[u] + value;
            value = h[u] * hz;
            pbuf[(0 * 32 + lane) * FDN2_PS + u] = o[u] * lw;
            pbuf[(1 * 32 + lane) * FDN2_PS + u] = o[u] * rw;
          }
        } else {
#pragma unroll 1
          for (int u = 0; u < gn; u++) {
            const float d = rb[g0 + u];
            rb[g0 + u] = din[g0 + u] + value;
            f0 = f1; f1 = f2; f2 = d;
            const float o = (w0 * f0 + w1 * f1) + w2 * f2;
            value = hadamard1(o) * hz;
            pbuf[(0 * 32 + lane) * FDN2_PS + u] = o * lw;
            pbuf[(1 * 32 + lane) * FDN2_PS + u] = o * rw;
          }
        }
        __syncwarp();
        if (lane < 2 * gn) {   // 16 lanes: (channel, sample) pairs; left fold over the 32 lines in index order
          const int ch = lane >= gn ? 1 : 0, u = lane - ch * gn, t = g0 + u;
          float s = pbuf[(ch * 32 + 0) * FDN2_PS + u];
#pragma unroll
          for (int l = 1; l < 32; l++) s += pbuf[(ch * 32 + l) * FDN2_PS + u];
          s = s * (ch ? c1 : c0);
          if (a.scalar_row >= 0) s = dbuf[(cur * 2 + ch) * 64 + t] + s * scalar;
          obuf[ch * 64 + t] = s;
          if (orow) orow[(size_t)ch * a.out_stride + t] = s;
        }
        __syncwarp();
      }
      if (tb == nb) {                                     // the warp that finished the block hands over to the next block:
        vbuf[lane] = value;                               //   feedback value entering its first sample
        float* hn = rbuf + ((cur ^ 1) * 32 + lane) * FDN2_RS;   // FIR history d[-3..-1] (the slice itself is overwritten in place)
        hn[0] = f0; hn[1] = f1; hn[2] = f2;
      }
    }
    named_bar(bar_id, bar_n);                             // all new samples are in place
    // write back this warp's lines and advance their ring indices
    if (active) {
#pragma unroll 4
      for (int l = l0; l < l1; l++) {
        const uint32_t L = tlen[l], base = toff[l], i0 = tidx[l];
#pragma unroll
        for (int q = 0; q < 2; q++) {
          const int t = lane + 32 * q;
          if (t < nb) {
            uint32_t pos = i0 + (uint32_t)t;
            pos -= (pos >= L) ? L : 0u;
            ring[base + pos] = rbuf[(cur * 32 + l) * FDN2_RS + 3 + t];
          }
        }
      }
      __syncwarp();
      if (lane >= l0 && lane < l1) { uint32_t i = tidx[lane] + (uint32_t)nb; const uint32_t L = tlen[lane]; i -= (i >= L) ? L : 0u; tidx[lane] = i; }
    }
    if (a.partial) {
      __syncthreads();
      for (int e = threadIdx.x; e < 2 * nb; e += blockDim.x) {
        const int ch = e / nb, t = e - ch * nb;
        const size_t ob = (size_t)(obuf - sm);
        float s = (blockIdx.x * VPB + 0 < a.V) ? fdn_smem[0 * (size_t)VF + ob + ch * 64 + t] : 0.0f;
        for (int w = 1; w < VPB; w++) s += (blockIdx.x * VPB + w < a.V) ? fdn_smem[(size_t)w * VF + ob + ch * 64 + t] : 0.0f;
        a.partial[((size_t)blockIdx.x * 2 + ch) * a.n + t0 + t] = s;
      }
      __syncthreads();
    } else {
      named_bar(bar_id, bar_n);
    }
    cur ^= 1;
  }
  cp_async_wait<0>();
  if (active && k == 0) {
    const uint32_t V = a.V;
    const float* h = rbuf + (cur * 32 + lane) * FDN2_RS;
    a.state[(size_t)(a.s0 + lane) * V + v] = __float_as_uint(vbuf[lane]);
    a.state[(size_t)(a.s0 + 32 + 4 * lane) * V + v] = tidx[lane];
    a.state[(size_t)(a.s0 + 32 + 4 * lane + 1) * V + v] = __float_as_uint(h[0]);
    a.state[(size_t)(a.s0 + 32 + 4 * lane + 2) * V + v] = __float_as_uint(h[1]);
    a.state[(size_t)(a.s0 + 32 + 4 * lane + 3) * V + v] = __float_as_uint(h[2]);
  }
}

}  // namespace fdsp
