// C++ host mirror smoke test (no GPU needed): graph expressions written like the reference's compile, get the
// reference's arities, type expressions and ping hashes; arity errors surface as exceptions; a Bank fails loudly
// without a CUDA device.
#include <cstdio>
#include <vector>

#include "fundsp_b200.hpp"

using namespace fundsp_b200;

// FNV-1a over the device program's parameter and state words: two builds of the same graph agree to the last bit of every coefficient
static unsigned long long words_hash(const An& g) {
  int np = 0, ns = 0, nu = 0;
  fdsp_node_lowering(g.get(), nullptr, 0, nullptr, 0, nullptr, 0, &np, &ns, &nu);
  std::vector<uint32_t> P(np + 1), S(ns + 1), U(nu + 1);
  fdsp_node_lowering(g.get(), P.data(), np, S.data(), ns, U.data(), nu, &np, &ns, &nu);
  unsigned long long h = 1469598103934665603ull;
  for (int i = 0; i < np; i++) { h ^= P[i]; h *= 1099511628211ull; }
  for (int i = 0; i < ns; i++) { h ^= S[i]; h *= 1099511628211ull; }
  return h;
}

int main() {
  An g = sine_hz(440.0f) >> lowpass_hz(1000.0f, 1.0f);
  std::printf("sig %s\n", g.signature().c_str());
  uint64_t h[8];
  int n = fdsp_node_leaf_hashes(const_cast<fdsp_node*>(g.get()), h, 8);
  for (int i = 0; i < n; i++) std::printf("hash %016llx\n", (unsigned long long)h[i]);
  // FM voice (README.md:102) and a filtered-noise bus, exercising precedence: * before + before >> before & before ^ before |
  float f = 220.0f, m = 2.0f;
  An fm = sine_hz(f) * f * m + f >> sine();
  std::printf("fm %d %d %s\n", fm.inputs(), fm.outputs(), fm.signature().c_str());
  An bus = (noise().seed(7) >> lowpass_hz(500.0f, 1.0f) & noise() >> highpass_hz(2000.0f, 1.0f)) | !zero() >> white() * 0.5f;
  std::printf("bus %d %d\n", bus.inputs(), bus.outputs());
  An rev = stacki(4, [](int i) { return delay(0.01 * (i + 1)) >> fir3(0.5f); });
  std::printf("stacki %d %d\n", rev.inputs(), rev.outputs());
  // the wider opcode set keeps the reference's names and argument order
  An synth = (poly_saw_hz(110.0f) & 0.5f * (dc(55.0f) >> dsf_saw_r(0.6f))) >> lowrez_hz(900.0f, 0.4f) >> shape(Tanh{1.5f}) >> dcblock()
             >> (pass() & 0.3f * feedback_unit(0.02, 0.5f * lowpole_hz(3000.0f))) >> pan(0.25f)
             >> (multipass(2) & 0.25f * reverb3_stereo(2.0, 0.5, lowpole_hz(8000.0f)));
  std::printf("synth %d %d %s\n", synth.inputs(), synth.outputs(), synth.signature().c_str());
  An misc = pink() | brown() | (noise() >> convolve({1.0f, 0.5f, 0.25f})) | (dc(220.0f) >> lorenz()) | ((noise() | dc(800.0f, 1.0f, 0.5f)) >> morph())
            | (noise() >> follow(0.01f)) | (var(0.5f) * mls()) | ((noise() | dc(0.004f)) >> tap(0.001f, 0.01f));
  std::printf("misc %d %d\n", misc.inputs(), misc.outputs());
  An nlb = (noise() >> dlowpass_hz(Tanh{1.0f}, 1200.0f, 2.0f)) | ((noise() | dc(900.0f, 1.5f, 2.0f)) >> fbell(Softsign{0.8f})) | (noise() >> fresonator_hz(Clip{1.0f}, 700.0f, 4.0f));
  std::printf("nlb %d %d %s\n", nlb.inputs(), nlb.outputs(), nlb.signature().c_str());
  An wide = ((dc(110.0f, 0.3f) >> pulse()) | (noise() >> phase_synth(2))) >> rotate(0.5f, 0.8f) >> mixer(2, 3, {0.5f, -0.25f, 0.125f, 1.0f, 1.0f, 1.0f});
  std::printf("wide %d %d %s %016llx\n", wide.inputs(), wide.outputs(), wide.signature().c_str(), words_hash(wide));
  std::vector<float> wave(64); for (int i = 0; i < 64; i++) wave[i] = (float)i / 64.0f - 0.5f;
  An smp = (dc(0.75f) >> resample(playwave(wave, 8))) | (playwave_at(wave, 4, 40) >> meter(Meter::Rms(0.05))) | (noise() >> meter(Meter::Peak(0.1)) >> limiter(0.003f, 0.02f));
  std::printf("smp %d %d %s %016llx\n", smp.inputs(), smp.outputs(), smp.signature().c_str(), words_hash(smp));
  An ev = event(saw_hz(220.0f) >> lowpass_hz(900.0f, 2.0f), 0.0125, 0.75, Fade::Power, 0.01, 0.2);
  std::printf("event %d %d %s %016llx\n", ev.inputs(), ev.outputs(), ev.signature().c_str(), words_hash(ev));
  An vib = lfo([](double t, double* o, void*) { o[0] = 220.0 + 10.0 * t; }, 1, nullptr, 0.05) >> sine();
  std::printf("vib %d %d %s\n", vib.inputs(), vib.outputs(), vib.signature().c_str());
  An r1 = reverb_stereo(12.0, 2.5, 0.4f), r4 = reverb4_stereo(20.0, 3.0);
  std::printf("reverb_stereo %s %016llx\n", r1.signature().c_str(), words_hash(r1));
  std::printf("reverb4_stereo %s %016llx\n", r4.signature().c_str(), words_hash(r4));
  try {
    An bad = pass() >> (pass() | pass());
    std::printf("arity NOT detected\n");
    return 1;
  } catch (const Error& e) { std::printf("arity error %d\n", e.code); }
  if (fdsp_device_count() == 0) {
    std::vector<An> voices;
    voices.push_back(saw_hz(110.0f) >> lowpass_hz(800.0f, 2.0f));
    try { Bank b(voices); std::printf("bank created without GPU?\n"); return 1; } catch (const Error& e) { std::printf("bank error: %s\n", e.what()); }
  }
  return 0;
}
