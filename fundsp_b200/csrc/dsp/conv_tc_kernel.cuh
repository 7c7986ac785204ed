// fundsp_b200 tensor-core path of `convolve(h)` (reference src/convolve.rs:9-59, ID 100; the same contraction as a long `Fir`,
// src/fir.rs:43-89): V voices share ONE impulse response h of K taps, so a block of their outputs is a dense GEMM against the
// Toeplitz matrix of h — the one place on this path where tensor cores apply (north-star: "tensor cores are used only on the dense
// FIR/convolve tap contraction as a batched GEMM").
//
//     Y[v, t0 + n] = sum_j X[v, t0 - P + j] * T[j, n],     T[j, n] = h[P + n - j]  (0 outside the band),  j < N + P,  P = K - 1 rounded up to 4
//
//   M = 128 voices per CTA tile, N = 128 output samples, contraction over the N + K - 1 input samples the tile can see, in chunks
//   of 32 floats (one 128-byte swizzle row). Flops per tile = 2 * 128 * 128 * (127 + K)  (2 * V * 64 * (63 + K) per
//   64-block, here with N = 128), of which 2 * 128 * 128 * K are the convolution's own.
//
// Precision: TF32 keeps 10 mantissa bits, the bar is 1e-5 of the output peak, so the product is taken in the 3xTF32 split
//     x = xh + xl,  h = hh + hl   (xh = x with the low 13 mantissa bits cleared — what the tensor core reads anyway —, xl = x - xh exactly)
//     x * h  ~  xh*hl + xl*hh + xh*hh        (the dropped xl*hl is <= 2^-20 |x h|)
// as three `tcgen05.mma.kind::tf32` per 8-wide k-step. A "hi" operand is simply the f32 array itself; only the "lo" arrays are
// materialised (conv_split_lo_kernel, conv_toeplitz_kernel).
// Accumulation: the tensor core rounds the FP32 accumulator at every MMA, and a 4096-tap response is 1560 MMAs in a row — measured,
// one accumulator gave 1.3e-5 of the peak at K = 4096 (error growing with K). So the sum is spread over FOUR accumulators in TMEM
// (all 512 columns): the xh*hh products of chunks c = 0, 1, 2 (mod 3) in three of them — a third of the sequential roundings each —
// and both cross terms, 2^-11 smaller, in the fourth, where their roundings do not count; the epilogue adds the four in FP32.
//
// Kernel (one output tile per CTA, 192 threads): warp 0 = TMA producer (4 `cp.async.bulk.tensor.2d` per stage: X, Xlo, T, Tlo tiles,
// 128-byte swizzle, 3-stage ring of 64 KB, full/empty mbarriers), warp 1 = TMEM allocation + MMA issue (one elected lane; smem
// stage released by `tcgen05.commit`), warps 2-5 = epilogue (`tcgen05.ld` 32x32b, rows straight to HBM as 16-byte stores).
// X rows live in HBM as [voice][H + chunk] with the last H >= K - 1 samples of the previous chunk in front (conv_history_kernel
// moves them there), so a window never leaves its row; rows past V and columns past the data are zero-filled by TMA / ignored
// by the epilogue (the Toeplitz band is causal: they only reach outputs that are not stored).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace fdsp {

constexpr int CTC_M = 128, CTC_N = 128, CTC_KC = 32, CTC_STAGES = 3;
constexpr int CTC_TILE_A = CTC_M * CTC_KC * 4, CTC_TILE_B = CTC_N * CTC_KC * 4;          // bytes: 16 KB each
constexpr int CTC_STAGE_BYTES = 2 * CTC_TILE_A + 2 * CTC_TILE_B;                          // X, Xlo, T, Tlo
constexpr int CTC_SMEM = CTC_STAGES * CTC_STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers + tmem address*/;

struct ConvTcArgs {
  float* y; uint32_t y_stride, y_offset;       // output rows y[row_map[v] * y_stride + y_offset + t]
  const uint32_t* row_map;
  uint32_t V, n;                               // voices, samples of this launch
  uint32_t K, H;                               // taps; history columns in front of every X row (multiple of 32, >= K - 1)
};

__device__ __forceinline__ uint32_t ctc_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ctc_mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void ctc_mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void ctc_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tCW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra CD_%=;\n\tbra CW_%=;\n\tCD_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void ctc_tma_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
// shared-memory matrix descriptor, K-major, 128-byte swizzle: start address >> 4, LBO (unused with swizzle) = 1, SBO = 1024 B between
// 8-row groups, version 1 (sm_100), layout type 2 = SWIZZLE_128B   (cute/arch/mma_sm100_desc.hpp SmemDescriptor)
__device__ __forceinline__ uint64_t ctc_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t CTC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(CTC_N >> 3) << 17) | ((uint32_t)(CTC_M >> 4) << 24);
__device__ __forceinline__ void ctc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(CTC_IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void ctc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// grid = (time tiles, voice tiles); block = 192. STEP < 4 builds bring-up probes (tests/cpp/conv_tc_probe.cu): 1 = TMEM allocation only,
// 2 = + TMA pipeline, 3 = + MMA issue, 4 = the kernel.
template <int STEP = 4>
__global__ void __launch_bounds__(192, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap mx, const __grid_constant__ CUtensorMap mxl,
                                                         const __grid_constant__ CUtensorMap mt, const __grid_constant__ CUtensorMap mtl, const ConvTcArgs a) {
  extern __shared__ uint8_t ctc_raw[];
  const uint32_t base = (ctc_smem(ctc_raw) + 1023u) & ~1023u;                  // 128-byte swizzle atoms are 1024-byte aligned
  const uint32_t bars = base + CTC_STAGES * CTC_STAGE_BYTES;                   // full[S], empty[S], accum, then the TMEM address word
  auto full_bar = [&](int s) { return bars + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bars + 8u * (uint32_t)(CTC_STAGES + s); };
  const uint32_t accum_bar = bars + 8u * (2 * CTC_STAGES);
  const uint32_t tmem_slot = bars + 8u * (2 * CTC_STAGES + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t t0 = blockIdx.x * CTC_N, v0 = blockIdx.y * CTC_M;
  const uint32_t P = (a.K - 1u + 3u) & ~3u;                                    // look-back padded to 4 samples: every window starts 16-byte aligned
  const int nchunk = (int)((CTC_N + P + CTC_KC - 1u) / CTC_KC);                // contraction chunks of 32 input samples

  if (threadIdx.x == 0) {
    for (int s = 0; s < CTC_STAGES; s++) { ctc_mbar_init(full_bar(s), 1); ctc_mbar_init(empty_bar(s), 1); }
    ctc_mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM: 128 lanes x 4 accumulators of 128 FP32 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(4 * CTC_N) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0 && STEP >= 2) {   // ---- TMA producer
      const int col0 = (int)(a.H - P + t0);                                    // window start inside the X rows
      for (int c = 0; c < nchunk; c++) {
        const int s = c % CTC_STAGES, use = c / CTC_STAGES;
        if (use > 0 && STEP >= 3) ctc_mbar_wait(empty_bar(s), (uint32_t)(use - 1) & 1u);
        if (STEP == 2 && c >= CTC_STAGES) break;
        const uint32_t st = base + (uint32_t)s * CTC_STAGE_BYTES;
        ctc_mbar_expect(full_bar(s), CTC_STAGE_BYTES);
        ctc_tma_2d(st, &mx, col0 + c * CTC_KC, (int)v0, full_bar(s));
        ctc_tma_2d(st + CTC_TILE_A, &mxl, col0 + c * CTC_KC, (int)v0, full_bar(s));
        ctc_tma_2d(st + 2 * CTC_TILE_A, &mt, c * CTC_KC, 0, full_bar(s));
        ctc_tma_2d(st + 2 * CTC_TILE_A + CTC_TILE_B, &mtl, c * CTC_KC, 0, full_bar(s));
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && STEP >= 2) {   // ---- MMA issue: per 8-wide k-step  acc += xh*hl + xl*hh + xh*hh
      for (int c = 0; c < nchunk; c++) {
        if (STEP == 2 && c >= CTC_STAGES) break;
        const int s = c % CTC_STAGES, use = c / CTC_STAGES;
        ctc_mbar_wait(full_bar(s), (uint32_t)use & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = base + (uint32_t)s * CTC_STAGE_BYTES;
        const uint64_t dxh = ctc_desc(st), dxl = ctc_desc(st + CTC_TILE_A), dth = ctc_desc(st + 2 * CTC_TILE_A), dtl = ctc_desc(st + 2 * CTC_TILE_A + CTC_TILE_B);
        if (STEP >= 3)
#pragma unroll
        for (int k = 0; k < CTC_KC / 8; k++) {
          const uint64_t adv = (uint64_t)((k * 32) >> 4);                      // 8 floats = 32 bytes further inside the swizzle row
          ctc_mma(tmem + (uint32_t)((c % 3) * CTC_N), dxh + adv, dth + adv, (c >= 3 || k != 0) ? 1u : 0u);   // xh*hh: accumulator c mod 3
          ctc_mma(tmem + 3u * CTC_N, dxh + adv, dtl + adv, (c | k) != 0 ? 1u : 0u);                            // cross terms: the fourth
          ctc_mma(tmem + 3u * CTC_N, dxl + adv, dth + adv, 1u);
        }
        if (STEP >= 3) ctc_commit(empty_bar(s));                               // the stage is free once these MMAs have read it
      }
      if (STEP >= 3) ctc_commit(accum_bar);                                    // accumulator complete
    }
  } else if (STEP >= 4) {
    // ---- epilogue: warp w owns TMEM lanes 32 * (w % 4) .. +31 = voices of the tile; 4 x 32 columns each
    ctc_mbar_wait(accum_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;
    const uint32_t v = v0 + (uint32_t)(q * 32 + lane);
    const bool rok = v < a.V;
    float* yrow = rok ? a.y + (size_t)__ldg(a.row_map + v) * a.y_stride + a.y_offset + t0 : nullptr;
    const bool vec_ok = ((a.y_stride | a.y_offset) & 3u) == 0u;
#pragma unroll 1
    for (int cb = 0; cb < CTC_N; cb += 32) {
      uint32_t r[32];
      float acc[32];
#pragma unroll 1
      for (int part = 0; part < 4; part++) {   // ((a0 + a1) + a2) + cross
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(part * CTC_N + cb);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                       "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
                       "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; j++) acc[j] = part == 0 ? __uint_as_float(r[j]) : acc[j] + __uint_as_float(r[j]);
      }
#pragma unroll
      for (int j = 0; j < 32; j++) r[j] = __float_as_uint(acc[j]);
      if (rok) {
        const uint32_t left = a.n > t0 + (uint32_t)cb ? a.n - t0 - (uint32_t)cb : 0u;   // valid samples from this column on
        if (vec_ok && left >= 32u) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<uint4*>(yrow + cb + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; j++) if ((uint32_t)j < left) yrow[cb + j] = __uint_as_float(r[j]);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(4 * CTC_N) : "memory");
}

// xl = x - (x with the low 13 mantissa bits cleared), over the new columns of every X row
static __global__ void conv_split_lo_kernel(const float* __restrict__ x, float* __restrict__ xl, uint32_t V, uint32_t row_stride, uint32_t col0, uint32_t n) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
  if (t >= n || v >= V) return;
  const size_t e = (size_t)v * row_stride + col0 + t;
  const float f = x[e];
  xl[e] = f - __uint_as_float(__float_as_uint(f) & 0xffffe000u);
}
// the last H samples of every row (columns [n, n + H)) move to the front (columns [0, H)): history for the next chunk. One CTA per
// (row, array); the row's H values go through shared memory because source and destination overlap when n < H.
static __global__ void conv_history_kernel(float* x, float* xl, uint32_t row_stride, uint32_t H, uint32_t n) {
  extern __shared__ float hs[];
  float* row = (blockIdx.y ? xl : x) + (size_t)blockIdx.x * row_stride;
  for (uint32_t i = threadIdx.x; i < H; i += blockDim.x) hs[i] = row[n + i];
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < H; i += blockDim.x) row[i] = hs[i];
}
// T[n][j] = h[P + n - j] inside the band, 0 outside (P = K - 1 rounded up to 4: window column j is input sample t0 - P + j); hi = the f32
// itself, lo = h - tf32(h). Rows n < 128, J columns (multiple of 32).
static __global__ void conv_toeplitz_kernel(const float* __restrict__ h, uint32_t K, float* th, float* tl, uint32_t J) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x, n = blockIdx.y;
  if (j >= J) return;
  const int k = (int)((K - 1u + 3u) & ~3u) + (int)n - (int)j;
  const float f = (k >= 0 && k < (int)K) ? h[k] : 0.0f;
  th[(size_t)n * J + j] = f;
  tl[(size_t)n * J + j] = f - __uint_as_float(__float_as_uint(f) & 0xffffe000u);
}

}  // namespace fdsp
