// Warp-specialised form of the voice-bank kernel for graphs whose top level is a Pipe: the front stage A and the back
// stage B of every voice run in DIFFERENT warps of the same CTA and hand 8-sample groups over through shared memory.
//
// Why: one voice is a serial recurrence, and a bank of V voices has only V/32 voice-warps — 512 for the 16 384-voice
// headline bank against 592 warp schedulers — so the plain kernel (bank_kernel.cuh) runs one warp per scheduler and is
// bound by that warp's own dependency latency. Splitting the Pipe gives every scheduler two warps with independent
// instruction streams (the oscillator of samples t+16.. overlaps the filter of samples t..), which is the only
// parallelism left once the voices are spoken for. Arithmetic per node is untouched (same `step`/`step8` code, same
// order), so results stay bit-identical to bank_kernel; nodes only communicate through their buffers in the reference
// too (src/audionode.rs:1445-1449 Pipe::process: X into a temp buffer, then Y).
//
// Layout: 2*NT threads; thread t < NT runs stage A of voice (blockIdx.x*NT + t), thread NT+t runs stage B of the same
// voice. Producer warp w and consumer warp w own a private ring of NSLOT hand-off slots [A::OUT][HS][32 voices] and two
// mbarriers per slot (full / empty, 32 arrivals each): warp pairs never wait for other pairs. The CTA mix tile is
// reduced by the consumer half alone (named barrier 1).
#pragma once
#include "bank_kernel.cuh"

namespace fdsp {

// ---- where to cut the Pipe: the top-level cut, or one re-association to either side (Pipe is associative and the
// depth-first word order of parameters/state is the same for every association), whichever balances the stages best
template <class X, class Y> struct CutCost { static constexpr int a = Cost<X>::value, b = Cost<Y>::value, worst = a > b ? a : b; };
template <class G> struct PipeParts { static constexpr bool is_pipe = false; typedef G A; typedef G B; };
template <class X, class Y> struct PipeParts<Pipe<X, Y>> { static constexpr bool is_pipe = true; typedef X A; typedef Y B; };
template <class P, class Q, class Y> struct PipeParts<Pipe<Pipe<P, Q>, Y>> {
  static constexpr bool is_pipe = true;
  static constexpr bool inner = CutCost<P, Pipe<Q, Y>>::worst < CutCost<Pipe<P, Q>, Y>::worst;
  template <bool I, class D = void> struct Pick { typedef Pipe<P, Q> A; typedef Y B; };
  template <class D> struct Pick<true, D> { typedef P A; typedef Pipe<Q, Y> B; };
  typedef typename Pick<inner>::A A; typedef typename Pick<inner>::B B;
};
template <class G> struct WsOk {
  typedef PipeParts<G> PP;
  static constexpr bool value = PP::is_pipe && GroupPlan<G>::ok && GroupPlan<G>::code <= FDSP_GROUP_COST && Cost<typename PP::A>::value >= 24 && Cost<typename PP::B>::value >= 12 &&
                                PP::A::OUT >= 1 && PP::A::OUT <= 2;
};
__host__ __device__ constexpr int ws_hand_samples(int outs) { return mix_tile_samples(outs) < 16 ? mix_tile_samples(outs) : 16; }
__host__ __device__ constexpr size_t ws_hand_floats(int mid, int outs, int nt, bool mix) { return (size_t)2 * mid * (mix ? ws_hand_samples(outs) : 16) * nt; }

FDSP_DEV void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
FDSP_DEV void bar_named(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

template <class G, int NT, int MODE, bool TB>
__global__ void __launch_bounds__(2 * NT, 1) bank_kernel_ws(const BankArgs a) {
  typedef typename PipeParts<G>::A A;
  typedef typename PipeParts<G>::B B;
  constexpr int IN = A::IN, MID = A::OUT, OUT = B::OUT;
  constexpr int TS = (MODE & 2) ? mix_tile_samples(OUT) : 64;
  constexpr int HS = (MODE & 2) ? ws_hand_samples(OUT) : 16;   // samples per hand-off slot; TS is a multiple of HS
  constexpr int NSLOT = 2, NW = NT / 32;
  static_assert(MID == B::IN, "Pipe arity");
  extern __shared__ __align__(16) float tile[];                // MODE&2: [OUT][TS][NT+1]; then hand-off rings; then tables
  float* hand = tile + ((MODE & 2) ? mix_tile_floats(OUT, NT) : 0);
  __shared__ __align__(8) unsigned long long bars[1 + 2 * NSLOT * NW];
  const uint32_t tid = threadIdx.x;
  const bool producer = tid < NT;
  const uint32_t lt = producer ? tid : tid - NT;               // voice slot inside the CTA
  const uint32_t w = lt >> 5, lane = lt & 31u;
  const uint32_t vpc = a.vpc ? a.vpc : (uint32_t)NT;
  const uint32_t v = blockIdx.x * vpc + lt;
  const bool active = lt < vpc && v < a.V;
  const uint32_t bar0 = smem_addr(&bars[0]);
  auto full_bar = [&](uint32_t slot) { return bar0 + 8u * (1u + (w * NSLOT + slot) * 2u); };
  auto empty_bar = [&](uint32_t slot) { return bar0 + 8u * (1u + (w * NSLOT + slot) * 2u + 1u); };
  if (tid == 0) {
    mbar_init(bar0, 1);
    for (int q = 1; q < 1 + 2 * NSLOT * NW; q++) mbar_init(bar0 + 8u * q, 32);
  }
  __syncthreads();

  CtxT<TB> c;
  c.tsm = 0u; c.tsm_kind = -1;
  if (TB) {
    constexpr int KIND = WaveKind<G>::value >= 0 ? WaveKind<G>::value : 0;
    float* tsm = hand + ws_hand_floats(MID, OUT, NT, (MODE & 2) != 0);
    const uint32_t bytes = (uint32_t)a.wt[KIND].total * 4u;
    if (tid == 0) {
      mbar_expect_tx(bar0, bytes);
      const char* src = reinterpret_cast<const char*>(a.wt[KIND].data);
      for (uint32_t o = 0; o < bytes; o += 32768u) bulk_g2s(smem_addr(tsm) + o, src + o, (bytes - o) < 32768u ? (bytes - o) : 32768u, bar0);
    }
    mbar_wait(bar0, 0);
    c.tsm = smem_addr(tsm); c.tsm_kind = KIND;
  }
  c.wt = a.wt; c.dl = a.dline; c.V = a.V; c.v = v; c.sr = a.sr; c.sd64 = a.sd64; c.sd32 = a.sd32;
  float* const ring = hand + (size_t)w * NSLOT * MID * HS * 32 + lane;   // this warp pair's slots: [slot][MID][HS][32]
  uint32_t it = 0;                                                      // hand-off counter (same sequence on both sides)

  if (producer) {
    typename A::R r;
    if (active) { Loader l{a.params, a.state, a.uniform, a.V, v, 0u, 0u, 0u, 0u}; A::load(r, l); }
#pragma unroll 1
    for (uint32_t t0 = 0; t0 < a.n; t0 += 64) {
      const int nb = (a.n - t0) < 64u ? (int)(a.n - t0) : 64;
      const int nfull = nb & ~7;
      c.n = nb;
      const float* irow = (IN > 0) ? a.in + a.in_offset + t0 : nullptr;
#pragma unroll 1
      for (int s0 = 0; s0 < nb; s0 += HS, it++) {
        const uint32_t slot = it % NSLOT, use = it / NSLOT;
        if (use > 0) mbar_wait(empty_bar(slot), (use - 1u) & 1u);        // stage B has drained this slot
        float* hs = ring + (size_t)slot * MID * HS * 32;
        const int s1 = (s0 + HS) < nb ? (s0 + HS) : nb;
        if (active) {
          const int gend = s1 < nfull ? s1 : nfull;
          c.rem = false;
#pragma unroll 1
          for (int g = s0; g < gend; g += 8) {
            Fr8<IN> in8; Fr8<MID> o8;
#pragma unroll
            for (int k = 0; k < IN; k++) {
#pragma unroll
              for (int j = 0; j < 8; j++) in8.v[k][j] = __ldg(irow + (size_t)k * a.in_stride + g + j);
            }
            c.i = g; c.first = true;
            group_step<A>(r, c, in8, o8);
#pragma unroll
            for (int k = 0; k < MID; k++) {
#pragma unroll
              for (int j = 0; j < 8; j++) hs[(k * HS + (g - s0) + j) * 32] = o8.v[k][j];
            }
          }
          if (s1 == nb) {  // end of the block: SIMD wrap-up, then the (size & 7) tail through the tick path
            A::end_simd(r);
            c.rem = true; c.first = false;
#pragma unroll 1
            for (int i = nfull; i < nb; i++) {
              Fr<IN> in; Fr<MID> o;
#pragma unroll
              for (int k = 0; k < IN; k++) in.v[k] = __ldg(irow + (size_t)k * a.in_stride + i);
              c.i = i;
              A::template step<false>(r, c, in, o);
#pragma unroll
              for (int k = 0; k < MID; k++) hs[(k * HS + (i - s0)) * 32] = o.v[k];
            }
          }
        }
        mbar_arrive(full_bar(slot));                                     // 32 arrivals: each lane releases its own stores
      }
    }
    if (active) { Saver s{a.state, a.V, v, 0u}; A::save(r, s); }
    return;
  }

  // ------------------------------------------------------------------ consumer half: stage B, outputs, CTA mix
  typename B::R r;
  if (active) {
    Loader l{a.params, a.state, a.uniform, a.V, v, 0u, 0u, 0u, 0u};
    typename A::R skip; A::load(skip, l);    // advances the word / delay-line cursors past stage A (its loads are dead code)
    B::load(r, l);
  } else if (MODE & 2) {
    for (int e = 0; e < OUT * TS; e++) tile[e * (NT + 1) + lt] = 0.0f;
  }
  const bool vec_ok = ((a.out_stride | a.out_offset) & 3u) == 0u;
#pragma unroll 1
  for (uint32_t t0 = 0; t0 < a.n; t0 += 64) {
    const int nb = (a.n - t0) < 64u ? (int)(a.n - t0) : 64;
    const int nfull = nb & ~7;
    c.n = nb;
    float* orow = (active && (MODE & 1)) ? a.out + (size_t)__ldg(a.row_map + v) * a.out_stride + a.out_offset + t0 : nullptr;
#pragma unroll 1
    for (int m0 = 0; m0 < nb; m0 += TS) {
      const int m1 = (m0 + TS) < nb ? (m0 + TS) : nb;
#pragma unroll 1
      for (int s0 = m0; s0 < m1; s0 += HS, it++) {
        const uint32_t slot = it % NSLOT, use = it / NSLOT;
        mbar_wait(full_bar(slot), use & 1u);
        const float* hs = ring + (size_t)slot * MID * HS * 32;
        const int s1 = (s0 + HS) < nb ? (s0 + HS) : nb;
        if (active) {
          const int gend = s1 < nfull ? s1 : nfull;
          c.rem = false;
#pragma unroll 1
          for (int g = s0; g < gend; g += 8) {
            Fr8<MID> in8; Fr8<OUT> o8;
#pragma unroll
            for (int k = 0; k < MID; k++) {
#pragma unroll
              for (int j = 0; j < 8; j++) in8.v[k][j] = hs[(k * HS + (g - s0) + j) * 32];
            }
            c.i = g; c.first = true;
            group_step<B>(r, c, in8, o8);
#pragma unroll
            for (int k = 0; k < OUT; k++) {
#pragma unroll
              for (int j = 0; j < 8; j++) {
                if (MODE & 2) tile[(k * TS + (g - m0) + j) * (NT + 1) + lt] = o8.v[k][j];
              }
              if (MODE & 1) {
                float* p = orow + (size_t)k * a.out_stride + g;
                if (vec_ok) {
                  *reinterpret_cast<float4*>(p) = make_float4(o8.v[k][0], o8.v[k][1], o8.v[k][2], o8.v[k][3]);
                  *reinterpret_cast<float4*>(p + 4) = make_float4(o8.v[k][4], o8.v[k][5], o8.v[k][6], o8.v[k][7]);
                } else {
#pragma unroll
                  for (int j = 0; j < 8; j++) p[j] = o8.v[k][j];
                }
              }
            }
          }
          if (s1 == nb) {
            B::end_simd(r);
            c.rem = true; c.first = false;
#pragma unroll 1
            for (int i = nfull; i < nb; i++) {
              Fr<MID> in; Fr<OUT> o;
#pragma unroll
              for (int k = 0; k < MID; k++) in.v[k] = hs[(k * HS + (i - s0)) * 32];
              c.i = i;
              B::template step<false>(r, c, in, o);
#pragma unroll
              for (int k = 0; k < OUT; k++) {
                if (MODE & 1) orow[(size_t)k * a.out_stride + i] = o.v[k];
                if (MODE & 2) tile[(k * TS + (i - m0)) * (NT + 1) + lt] = o.v[k];
              }
            }
          }
        }
        mbar_arrive(empty_bar(slot));
      }
      if (MODE & 2) {
        // CTA partial mix by the consumer half (same association as bank_kernel: two threads per row, four accumulators)
        bar_named(1, NT);
        constexpr int HALF = NT / 2, QN = HALF / 4, ROWS = OUT * TS;
        const int h = (int)(lt & 1u);
        const int c0 = h * HALF + (h * QN) % HALF, c1 = h * HALF + (QN + h * QN) % HALF, c2 = h * HALF + (2 * QN + h * QN) % HALF, c3 = h * HALF + (3 * QN + h * QN) % HALF;
#pragma unroll 1
        for (int eb = (int)(lt >> 5) * 16; eb < ROWS; eb += HALF) {
          const int e = eb + (int)((lt & 31u) >> 1);
          const bool ok = e < ROWS;
          const int k = e / TS, i = e - k * TS;
          const float* row = tile + (ok ? e : 0) * (NT + 1);
          float a0 = row[c0], a1 = row[c1], a2 = row[c2], a3 = row[c3];
#pragma unroll
          for (int q = 1; q < QN; q++) { a0 += row[c0 + q]; a1 += row[c1 + q]; a2 += row[c2 + q]; a3 += row[c3 + q]; }
          const float s = (a0 + a1) + (a2 + a3);
          const float other = __shfl_xor_sync(0xffffffffu, s, 1);
          if (ok && h == 0 && m0 + i < m1) a.partial[((size_t)blockIdx.x * OUT + k) * a.n + t0 + m0 + i] = s + other;
        }
        bar_named(1, NT);
      }
    }
  }
  if (active) { Saver s{a.state, a.V, v, (uint32_t)A::NS}; B::save(r, s); }
}

}  // namespace fdsp
