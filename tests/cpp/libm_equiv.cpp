// TEST INFRASTRUCTURE: the PRODUCT's scalar math (fundsp_b200/csrc/dsp/libm.cuh, compiled for the host through the FDSP_HOST_EMUL shims)
// against the ORACLE's independent restatement (oracle/fo_libm.h), bit for bit, over float bit patterns.
//   g++ -std=c++17 -O2 -ffp-contract=off -pthread tests/cpp/libm_equiv.cpp -o libm_equiv ; ./libm_equiv STRIDE   (STRIDE 1 = all 2^32 patterns)
// Functions: tanhf (the value on the Moog ladder's per-sample recurrence), sinf, cosf, tanf, expm1f, expf. NaNs compare as a class.
#define FDSP_HOST_EMUL 1
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "../../fundsp_b200/csrc/dsp/libm.cuh"
#include "../../oracle/fo_libm.h"

static inline float fromb_(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t bits_(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline bool same(float a, float b) { return (a != a && b != b) || bits_(a) == bits_(b); }

int main(int argc, char** argv) {
  const uint64_t stride = argc > 1 ? strtoull(argv[1], nullptr, 10) : 257;
  const unsigned nt = std::max(1u, std::thread::hardware_concurrency());
  std::vector<uint64_t> bad(nt * 6, 0), first(nt * 6, 0);
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; t++) th.emplace_back([&, t] {
    for (uint64_t u = t * stride; u < (1ull << 32); u += (uint64_t)nt * stride) {
      const float x = fromb_((uint32_t)u);
      const float p[6] = {fdsp::m::tanhf_(x), fdsp::m::sinf_(x), fdsp::m::cosf_(x), fdsp::m::tanf_(x), fdsp::m::expm1f_(x), fdsp::m::expf_(x)};
      const float o[6] = {fo::m::tanhf_(x), fo::m::sinf_(x), fo::m::cosf_(x), fo::m::tanf_(x), fo::m::expm1f_(x), fo::m::expf_(x)};
      for (int k = 0; k < 6; k++) if (!same(p[k], o[k])) { if (!bad[t * 6 + k]) first[t * 6 + k] = u; bad[t * 6 + k]++; }
    }
  });
  for (auto& x : th) x.join();
  const char* names[6] = {"tanhf", "sinf", "cosf", "tanf", "expm1f", "expf"};
  int rc = 0;
  for (int k = 0; k < 6; k++) {
    uint64_t b = 0, f = 0;
    for (unsigned t = 0; t < nt; t++) { b += bad[t * 6 + k]; if (bad[t * 6 + k] && !f) f = first[t * 6 + k]; }
    printf("%s: %llu mismatches%s\n", names[k], (unsigned long long)b, b ? "" : " (bit-identical)");
    if (b) { printf("  first at bits 0x%08llx\n", (unsigned long long)f); rc = 1; }
  }
  return rc;
}
