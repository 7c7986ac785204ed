// fundsp_b200 voice-bank kernel: one thread = one voice instance of the graph type G, V voices in
// lockstep. Replaces the reference's per-voice `AudioUnit::process` recursion (src/audiounit.rs:393-395,
// src/audionode.rs:85-105,1445-1449) driven by `Wave::render` (src/wave.rs:441-466).
//
// Launch shape: grid = ceil(V / NT) CTAs of NT threads; each thread keeps G::R (parameters + state) in
// registers for all blocks of the launch and walks time in 64-sample blocks (8 SIMD groups of 8 + tail),
// exactly the block structure the reference's `process` path exposes to the nodes.
// Outputs: MODE bit 0 = per-voice rows out[(v*OUT+c)*stride + off + t] written as coalesced-by-row
// 16-byte stores; MODE bit 1 = per-CTA index-order partial mix (shared-memory transpose + sequential
// column sums), finished by mix_reduce_kernel in CTA order (deterministic).
#pragma once
#include "nodes.cuh"
#include "bank_args.h"

namespace fdsp {



// ---- TMA (bulk async copy) + mbarrier primitives used to stage wavetables into shared memory
FDSP_DEV uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
FDSP_DEV void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
FDSP_DEV void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
FDSP_DEV void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
FDSP_DEV void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
}

// TB: stage the wavetable data of kind WaveKind<G> in shared memory behind the mix tile (one TMA bulk copy per
// 32 KB slice, single mbarrier), so the per-sample table taps become conflict-light LDS instead of divergent LDG.
#ifndef FDSP_MIN_CTAS
#define FDSP_MIN_CTAS 1   // one resident CTA per SM is all a bank needs: lets ptxas spend registers on overlapping the 8-sample group
#endif
#ifndef FDSP_NO_GROUP
#define FDSP_NO_GROUP 0
#endif
#ifndef FDSP_GROUP_COST
#define FDSP_GROUP_COST 256   // programs whose group form unrolls to at most this many instructions per sample are evaluated 8 samples at a time
#endif

template <class G, int NT, int MODE, bool TB>
__global__ void __launch_bounds__(NT, FDSP_MIN_CTAS) bank_kernel(const BankArgs a) {
  extern __shared__ __align__(16) float tile[];  // MODE&2: [OUT][TS][NT+1]; TB: table data after it
  const uint32_t tid = threadIdx.x;
  const uint32_t vpc = a.vpc ? a.vpc : (uint32_t)NT;
  const uint32_t v = blockIdx.x * vpc + tid;
  const bool active = tid < vpc && v < a.V;
  constexpr int IN = G::IN, OUT = G::OUT;
  constexpr int TS = (MODE & 2) ? mix_tile_samples(OUT) : 64;
  // big programs (e.g. the 32-line FDN in thread-per-voice form) are not unrolled over the 8-sample group
  constexpr bool GROUP = GroupPlan<G>::ok && GroupPlan<G>::code <= FDSP_GROUP_COST && !FDSP_NO_GROUP;
  constexpr int UNROLL = GROUP ? 8 : (Cost<G>::value <= 160 ? 8 : (Cost<G>::value <= 320 ? 4 : (Cost<G>::value <= 640 ? 2 : 1)));

  typename G::R r;
  CtxT<TB> c;
  c.tsm = 0u; c.tsm_kind = -1;
  if (TB) {
    constexpr int KIND = WaveKind<G>::value >= 0 ? WaveKind<G>::value : 0;
    __shared__ __align__(8) unsigned long long mbar;
    float* tsm = tile + ((MODE & 2) ? mix_tile_floats(OUT, NT) : 0);
    const uint32_t bytes = (uint32_t)a.wt[KIND].total * 4u;
    const uint32_t bar = smem_addr(&mbar);
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(bar, bytes);
      const char* src = reinterpret_cast<const char*>(a.wt[KIND].data);
      for (uint32_t o = 0; o < bytes; o += 32768u) bulk_g2s(smem_addr(tsm) + o, src + o, (bytes - o) < 32768u ? (bytes - o) : 32768u, bar);
    }
    mbar_wait(bar, 0);
    c.tsm = smem_addr(tsm); c.tsm_kind = KIND;
  }
  c.wt = a.wt; c.dl = a.dline; c.V = a.V; c.v = v; c.sr = a.sr; c.sd64 = a.sd64; c.sd32 = a.sd32;
  c.rp = a.params; c.rs0 = a.state0; c.ru = a.uniform; c.dl_total = a.dl_floats;
  if (active) {
    Loader l{a.params, a.state, a.uniform, a.V, v, 0u, 0u, 0u, 0u};
    G::load(r, l);
  } else if (MODE & 2) {
    for (int e = 0; e < OUT * TS; e++) tile[e * (NT + 1) + tid] = 0.0f;  // columns of absent voices stay zero
  }
  const bool vec_ok = ((a.out_stride | a.out_offset) & 3u) == 0u;

#pragma unroll 1
  for (uint32_t t0 = 0; t0 < a.n; t0 += 64) {
    const int nb = (a.n - t0) < 64u ? (int)(a.n - t0) : 64;
    const int nfull = nb & ~7;
    c.n = nb;
    float* orow = (active && (MODE & 1)) ? a.out + (size_t)__ldg(a.row_map + v) * a.out_stride + a.out_offset + t0 : nullptr;
    const float* irow = (IN > 0) ? a.in + a.in_offset + t0 : nullptr;
#pragma unroll 1
    for (int s0 = 0; s0 < nb; s0 += TS) {   // one pass when there is no mix tile (TS = 64)
      const int s1 = (s0 + TS) < nb ? (s0 + TS) : nb;
      if (active) {
        const int gend = s1 < nfull ? s1 : nfull;
        c.rem = false;
#pragma unroll 1
        for (int g = s0; g < gend; g += 8) {
          float ob[OUT > 0 ? OUT : 1][8];
          if constexpr (GROUP) {
            // small programs: node by node over the 8-sample group (group_step), all intermediates in registers
            Fr8<IN> in8; Fr8<OUT> o8;
#pragma unroll
            for (int k = 0; k < IN; k++) {
#pragma unroll
              for (int j = 0; j < 8; j++) in8.v[k][j] = __ldg(irow + (size_t)k * a.in_stride + g + j);
            }
            c.i = g; c.first = true;
            group_step<G>(r, c, in8, o8);
#pragma unroll
            for (int k = 0; k < OUT; k++) {
#pragma unroll
              for (int j = 0; j < 8; j++) {
                ob[k][j] = o8.v[k][j];
                if (MODE & 2) tile[(k * TS + (g - s0) + j) * (NT + 1) + tid] = o8.v[k][j];
              }
            }
          } else {
#pragma unroll(UNROLL)
            for (int j = 0; j < 8; j++) {
              Fr<IN> in; Fr<OUT> o;
#pragma unroll
              for (int k = 0; k < IN; k++) in.v[k] = __ldg(irow + (size_t)k * a.in_stride + g + j);
              c.i = g + j; c.first = (j == 0);
              G::template step<false>(r, c, in, o);
#pragma unroll
              for (int k = 0; k < OUT; k++) {
                ob[k][j] = o.v[k];
                if (MODE & 2) tile[(k * TS + (g - s0) + j) * (NT + 1) + tid] = o.v[k];
              }
            }
          }
          if (MODE & 1) {
#pragma unroll
            for (int k = 0; k < OUT; k++) {
              float* p = orow + (size_t)k * a.out_stride + g;
              if (vec_ok) {
                *reinterpret_cast<float4*>(p) = make_float4(ob[k][0], ob[k][1], ob[k][2], ob[k][3]);
                *reinterpret_cast<float4*>(p + 4) = make_float4(ob[k][4], ob[k][5], ob[k][6], ob[k][7]);
              } else {
#pragma unroll
                for (int j = 0; j < 8; j++) p[j] = ob[k][j];
              }
            }
          }
        }
        if (s1 == nb) {  // end of the block: wrap up the SIMD part, then the (size & 7) tail through the tick path
          G::end_simd(r);
          c.rem = true; c.first = false;
#pragma unroll 1
          for (int i = nfull; i < nb; i++) {
            Fr<IN> in; Fr<OUT> o;
#pragma unroll
            for (int k = 0; k < IN; k++) in.v[k] = __ldg(irow + (size_t)k * a.in_stride + i);
            c.i = i;
            G::template step<false>(r, c, in, o);
#pragma unroll
            for (int k = 0; k < OUT; k++) {
              if (MODE & 1) orow[(size_t)k * a.out_stride + i] = o.v[k];
              if (MODE & 2) tile[(k * TS + (i - s0)) * (NT + 1) + tid] = o.v[k];
            }
          }
        }
      }
      if (MODE & 2) {
        // CTA partial mix: two threads per (channel, sample) row, each folding one half of the voice columns with four
        // interleaved accumulators; fixed association ((a0+a1)+(a2+a3)) then low half + high half -> deterministic.
        __syncthreads();
        constexpr int HALF = NT / 2, QN = HALF / 4, ROWS = OUT * TS;
        const int h = (int)(tid & 1u);
        // the odd thread starts QN columns further so that the 32 lanes of a warp (16 rows x 2 halves) hit 32 banks
        const int c0 = h * HALF + (h * QN) % HALF, c1 = h * HALF + (QN + h * QN) % HALF, c2 = h * HALF + (2 * QN + h * QN) % HALF, c3 = h * HALF + (3 * QN + h * QN) % HALF;
#pragma unroll 1
        for (int eb = (int)(tid >> 5) * 16; eb < ROWS; eb += HALF) {   // warp-uniform trip count (full-mask shuffle below)
          const int e = eb + (int)((tid & 31u) >> 1);
          const bool ok = e < ROWS;
          const int k = e / TS, i = e - k * TS;
          const float* row = tile + (ok ? e : 0) * (NT + 1);
          float a0 = row[c0], a1 = row[c1], a2 = row[c2], a3 = row[c3];
#pragma unroll
          for (int q = 1; q < QN; q++) { a0 += row[c0 + q]; a1 += row[c1 + q]; a2 += row[c2 + q]; a3 += row[c3 + q]; }
          const float s = (a0 + a1) + (a2 + a3);
          const float other = __shfl_xor_sync(0xffffffffu, s, 1);
          if (ok && h == 0 && s0 + i < s1) a.partial[((size_t)blockIdx.x * OUT + k) * a.n + t0 + s0 + i] = s + other;
        }
        __syncthreads();
      }
    }
  }
  if (active) {
    Saver s{a.state, a.V, v, 0u};
    G::save(r, s);
  }
  if ((MODE & 2) && a.ticket) {   // fused finish of the mix-down: same CTA-order left fold as mix_reduce_kernel
    __shared__ uint32_t s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1u) ? 1u : 0u;
    __syncthreads();
    if (s_last) {
      __threadfence();
      for (uint32_t e = tid; e < (uint32_t)OUT * a.n; e += NT) {
        const uint32_t ch = e / a.n, t = e - ch * a.n;
        float s = __ldcg(a.partial + (size_t)ch * a.n + t);
        for (uint32_t b = 1; b < gridDim.x; b++) s += __ldcg(a.partial + ((size_t)b * OUT + ch) * a.n + t);
        float* p = a.mix + (size_t)ch * a.mix_stride + a.mix_offset + t;
        *p = a.mix_accumulate ? *p + s : s;
      }
      if (tid == 0) *a.ticket = 0u;
    }
  }
}

// Finishes the mix-down: mix[c][off + t] (+)= sum over CTAs in CTA order (deterministic).
static __global__ void mix_reduce_kernel(const float* partial, uint32_t nparts, uint32_t outs, uint32_t n, float* mix,
                                  uint32_t mix_stride, uint32_t mix_offset, int accumulate) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= outs * n) return;
  const uint32_t c = e / n, t = e % n;
  float s = partial[(size_t)c * n + t];
  for (uint32_t b = 1; b < nparts; b++) s += partial[((size_t)b * outs + c) * n + t];
  float* p = mix + (size_t)c * mix_stride + mix_offset + t;
  *p = accumulate ? *p + s : s;
}


// Mix-down of per-voice rows in a FIXED association order, for parity with a reference `Net` whose outputs are adder
// trees (src/net.rs:1693-1767 Net::bus adds one Binop<Add,Pass,Pass> vertex per output per call):
//   pairwise = level-wise adjacent pairing with the odd element carried up (the balanced `Net::bus` tree),
//   chain    = left fold in voice order (a left-leaning chain of `&`, or the index-order sum of a Sequencer).
// One thread per (channel, sample); rows are read coalesced along time; the association is a binary-carry stack.
// Level 1 of the balanced tree for big banks: the subtree sum of every COMPLETE block of 2^LB consecutive voices, one thread per (block,
// channel, sample). A complete aligned block is a whole subtree of the level-wise pairing, so its sum does not depend on the rest of the bank;
// tree_mix_kernel then pushes the block sums at level LB of its carry stack (exactly what pushing the block's voices one by one would have left
// there) and goes on with the tail voices. Same association, every bit — with V / 2^LB times the parallelism over the voices.
template <int LB>
static __global__ void tree_mix_block_kernel(const float* __restrict__ rows, uint32_t outs, uint32_t row_stride, uint32_t row_offset, uint32_t n, float* __restrict__ partial) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x, blk = blockIdx.y;
  if (e >= outs * n) return;
  const uint32_t ch = e / n, t = e - ch * n;
  const size_t vstep = (size_t)outs * row_stride;
  const float* p = rows + (size_t)ch * row_stride + row_offset + t + (size_t)blk * ((size_t)1 << LB) * vstep;
  float st[LB + 1];
  for (uint32_t v0 = 0; v0 < (1u << LB); v0 += 8) {
    float b[8];
#pragma unroll
    for (int u = 0; u < 8; u++) b[u] = __ldg(p + (size_t)(v0 + u) * vstep);
#pragma unroll
    for (int u = 0; u < 8; u++) {
      float y = b[u]; uint32_t q = v0 + u; int lvl = 0;
      while (q & 1u) { y = st[lvl] + y; q >>= 1; lvl++; }
      st[lvl] = y;
    }
  }
  partial[((size_t)blk * outs + ch) * n + t] = st[LB];
}

static __global__ void tree_mix_kernel(const float* __restrict__ rows, uint32_t V, uint32_t outs, uint32_t row_stride, uint32_t row_offset, uint32_t n,
                                       float* mix, uint32_t mix_stride, uint32_t mix_offset, int pairwise, const float* __restrict__ partial = nullptr,
                                       uint32_t nfull = 0, int lb = 0) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= outs * n) return;
  const uint32_t ch = e / n, t = e - ch * n;
  const float* p = rows + (size_t)ch * row_stride + row_offset + t;
  const size_t vstep = (size_t)outs * row_stride;
  float x;
  if (pairwise) {
    float st[32];
    // complete blocks of 2^lb voices arrive as their subtree sums (tree_mix_block_kernel): pushed at level lb with the block index as the carry pattern
    for (uint32_t bk = 0; bk < nfull; bk++) {
      float y = __ldg(partial + ((size_t)bk * outs + ch) * n + t); uint32_t q = bk; int lvl = lb;
      while (q & 1u) { y = st[lvl] + y; q >>= 1; lvl++; }
      st[lvl] = y;
    }
    for (uint32_t v0 = nfull << lb; v0 < V; v0 += 8) {
      float b[8];
#pragma unroll
      for (int u = 0; u < 8; u++) b[u] = (v0 + u < V) ? __ldg(p + (size_t)(v0 + u) * vstep) : 0.0f;
#pragma unroll
      for (int u = 0; u < 8; u++) {
        if (v0 + u < V) {
          float y = b[u]; uint32_t q = v0 + u; int lvl = 0;
          while (q & 1u) { y = st[lvl] + y; q >>= 1; lvl++; }
          st[lvl] = y;
        }
      }
    }
    int lvl = 0; uint32_t q = V;
    while (!(q & 1u)) { q >>= 1; lvl++; }
    x = st[lvl]; q >>= 1; lvl++;
    for (; q; q >>= 1, lvl++) if (q & 1u) x = st[lvl] + x;
  } else {
    x = __ldg(p);
    for (uint32_t v = 1; v < V; v++) x += __ldg(p + (size_t)v * vstep);
  }
  mix[(size_t)ch * mix_stride + mix_offset + t] = x;
}

}  // namespace fdsp
