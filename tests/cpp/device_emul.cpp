// TEST INFRASTRUCTURE — host emulation of ONE voice of a fused device program.
//
// Compiles the DEVICE node library (fundsp_b200/csrc/dsp/nodes.cuh) for the CPU (FDSP_HOST_EMUL shims in math.cuh) and walks it
// through the same block structure as bank_kernel (8-sample groups, end_simd, tick-path tail), so that the CPU-only test suite
// can compare the device templates against the oracle bit for bit. It is never linked into the product and is not a fallback:
// the product library has no CPU DSP path. Build: g++ -std=c++17 -O1 -ffp-contract=off -DGRAPH='<type expression>' device_emul.cpp
//
// stdin-free protocol: argv[1] = input blob, argv[2] = output file.
//   blob: u32 np, ns, nu, nin, n ; f64 sr ; u32 P[np], S[ns], U[nu] ; f32 in[nin][n] ;
//         u32 ntables ; per table: u32 kind, n, total ; f32 pitch[n] ; i32 off[n] ; i32 len[n] ; f32 data[total]   (guard-sample layout)
//   out : f32 y[OUT][n]
#define FDSP_HOST_EMUL 1
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../fundsp_b200/csrc/dsp/nodes.cuh"

using namespace fdsp;
typedef GRAPH G;

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 3;
  uint32_t hdr[5]; double sr;
  if (fread(hdr, 4, 5, f) != 5 || fread(&sr, 8, 1, f) != 1) return 4;
  const uint32_t np = hdr[0], ns = hdr[1], nu = hdr[2], nin = hdr[3], n = hdr[4];
  if ((int)nin != G::IN) { fprintf(stderr, "inputs %u != %d\n", nin, G::IN); return 5; }
  std::vector<uint32_t> P(np + 1), S(ns + 1), U(nu + 1);
  if (fread(P.data(), 4, np, f) != np || fread(S.data(), 4, ns, f) != ns || fread(U.data(), 4, nu, f) != nu) return 6;
  std::vector<float> in((size_t)(nin ? nin : 1) * n), out((size_t)(G::OUT ? G::OUT : 1) * n, 0.0f);
  if (nin && fread(in.data(), 4, (size_t)nin * n, f) != (size_t)nin * n) return 7;
  WaveTableDev wt[6] = {};
  std::vector<std::vector<float>> tdata;
  uint32_t ntab = 0;
  if (fread(&ntab, 4, 1, f) == 1) {
    tdata.reserve(ntab);   // keep the data pointers stable
    for (uint32_t t = 0; t < ntab; t++) {
      uint32_t h3[3];
      if (fread(h3, 4, 3, f) != 3 || h3[0] > 5 || h3[1] > 48) return 10;
      WaveTableDev& w = wt[h3[0]];
      w.n = (int)h3[1]; w.total = (int)h3[2];
      if (fread(w.pitch, 4, h3[1], f) != h3[1] || fread(w.off, 4, h3[1], f) != h3[1] || fread(w.len, 4, h3[1], f) != h3[1]) return 11;
      tdata.emplace_back(h3[2]);
      if (fread(tdata.back().data(), 4, h3[2], f) != h3[2]) return 12;
      w.data = tdata.back().data();
    }
  }
  fclose(f);

  typename G::R r;
  CtxT<false> c;
  c.wt = wt; c.tsm = 0u; c.tsm_kind = -1; c.V = 1; c.v = 0;
  c.sr = (float)sr; c.sd64 = (float)(1.0 / sr); c.sd32 = 1.0f / (float)sr;
  Loader l{P.data(), S.data(), U.data(), 1u, 0u, 0u, 0u, 0u, 0u};
  G::load(r, l);
  if (l.pi > np || l.si > ns || l.ui > nu) { fprintf(stderr, "word layout mismatch: consumed %u/%u/%u of %u/%u/%u\n", l.pi, l.si, l.ui, np, ns, nu); return 8; }
  std::vector<float> dline((size_t)l.dl + 1, 0.0f);
  fprintf(stderr, "dl=%u\n", l.dl);   // the delay-line floats the device program claims (the host must allocate exactly this)
  c.dl = dline.data();
  constexpr int IN = G::IN, OUT = G::OUT;
  constexpr bool GROUP = GroupPlan<G>::ok && GroupPlan<G>::code <= 256;   // bank_kernel's FDSP_GROUP_COST
  for (uint32_t t0 = 0; t0 < n; t0 += 64) {
    const int nb = (n - t0) < 64u ? (int)(n - t0) : 64;
    const int nfull = nb & ~7;
    c.n = nb; c.rem = false;
    for (int g = 0; g < nfull; g += 8) {
      if (GROUP) {
        Fr8<IN> in8; Fr8<OUT> o8;
        for (int k = 0; k < IN; k++) for (int j = 0; j < 8; j++) in8.v[k][j] = in[(size_t)k * n + t0 + g + j];
        c.i = g; c.first = true;
        group_step<G>(r, c, in8, o8);
        for (int k = 0; k < OUT; k++) for (int j = 0; j < 8; j++) out[(size_t)k * n + t0 + g + j] = o8.v[k][j];
      } else {
        for (int j = 0; j < 8; j++) {
          Fr<IN> a; Fr<OUT> b;
          for (int k = 0; k < IN; k++) a.v[k] = in[(size_t)k * n + t0 + g + j];
          c.i = g + j; c.first = (j == 0);
          G::template step<false>(r, c, a, b);
          for (int k = 0; k < OUT; k++) out[(size_t)k * n + t0 + g + j] = b.v[k];
        }
      }
    }
    G::end_simd(r);
    c.rem = true; c.first = false;
    for (int i = nfull; i < nb; i++) {
      Fr<IN> a; Fr<OUT> b;
      for (int k = 0; k < IN; k++) a.v[k] = in[(size_t)k * n + t0 + i];
      c.i = i;
      G::template step<false>(r, c, a, b);
      for (int k = 0; k < OUT; k++) out[(size_t)k * n + t0 + i] = b.v[k];
    }
  }
  FILE* o = fopen(argv[2], "wb");
  if (!o) return 9;
  fwrite(out.data(), 4, (size_t)OUT * n, o);
  fclose(o);
  return 0;
}
