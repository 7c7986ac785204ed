// fundsp_b200 host wavetable builder: bandlimited tables for saw/square/triangle/organ/soft_saw/hammond.
// Follows the reference's table construction (src/wavetable.rs:40-123 make_wave / Wavetable::new and the
// six generators at :493-623): pitches 20 Hz * 2^(k/4) up to 20 kHz, harmonics floor(22000/pitch) faded
// 20->22 kHz by smooth5, length clamp(32, 8192, next_pow2(4*harmonics)), global peak normalisation.
// The reference inverse-transforms with the `microfft` crate in f32; here an in-place radix-2 inverse FFT
// in f64 is used and the result rounded once to f32 (tables are read-only inputs to the GPU kernels).
#include "graph.h"

#include <complex>
#include <mutex>

namespace fdsp {
namespace host {
namespace {

typedef std::complex<double> cd;

void ifft_inplace(std::vector<cd>& a) {  // unnormalised inverse DFT: x[n] = sum_k a[k] e^{+2 pi i k n / N}
  const size_t n = a.size();
  for (size_t i = 1, j = 0; i < n; i++) {
    size_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) std::swap(a[i], a[j]);
  }
  for (size_t len = 2; len <= n; len <<= 1) {
    const double ang = 6.283185307179586476925 / (double)len;
    for (size_t i = 0; i < n; i += len)
      for (size_t k = 0; k < len / 2; k++) {
        cd w(cos(ang * (double)k), sin(ang * (double)k));
        cd u = a[i + k], v = a[i + k + len / 2] * w;
        a[i + k] = u + v; a[i + k + len / 2] = u - v;
      }
  }
}

double smooth5(double x) { return ((x * 6.0 - 15.0) * x + 10.0) * x * x * x; }
double clamp01(double x) { return fmin(fmax(x, 0.0), 1.0); }

typedef double (*PhaseFn)(uint32_t);
typedef double (*AmpFn)(double, uint32_t);

std::vector<float> make_wave(double pitch, PhaseFn phase, AmpFn amp) {  // src/wavetable.rs:40-79
  const double MAX_F = 22000.0, FADE_F = 20000.0;
  size_t harmonics = (size_t)floor(MAX_F / pitch);
  size_t target = 4 * harmonics, p2 = 1;
  while (p2 < target) p2 <<= 1;
  size_t length = std::min<size_t>(std::max<size_t>(p2, 32), 8192);
  std::vector<cd> a(length, cd(0.0, 0.0));
  for (size_t i = 1; i <= harmonics && i < length; i++) {
    double f = pitch * (double)i;
    double w = amp(pitch, (uint32_t)i) * smooth5(clamp01((f - MAX_F) / (FADE_F - MAX_F)));
    if (w > 0.0) {
      float r = (float)w, th = (float)(6.283185307179586 * phase((uint32_t)i));  // Complex32::from_polar
      a[i] = cd((double)(r * cosf(th)), (double)(r * sinf(th)));
    }
  }
  ifft_inplace(a);
  std::vector<float> out(length);
  for (size_t n = 0; n < length; n++) out[n] = (float)a[n].imag();
  return out;
}

double ph_saw(uint32_t i) { return (i & 1) == 1 ? 0.0 : 0.5; }
double am_saw(double, uint32_t i) { return 1.0 / (double)i; }
double ph_zero(uint32_t) { return 0.0; }
double am_square(double, uint32_t i) { return (i & 1) == 1 ? 1.0 / (double)i : 0.0; }
double ph_tri(uint32_t i) { return (i & 3) == 3 ? 0.5 : 0.0; }
double am_tri(double, uint32_t i) { return (i & 1) == 1 ? 1.0 / (double)(i * i) : 0.0; }
double ph_organ(uint32_t i) { return (i & 3) == 3 ? 0.5 : ((i & 1) == 1 ? 0.0 : 0.5); }
double am_organ(double, uint32_t i) { uint32_t z = (uint32_t)__builtin_ctz(i), j = i >> z; return 1.0 / (double)(i + j * j * j); }
double am_softsaw(double, uint32_t i) { return 1.0 / (double)(i * i); }
double am_hammond(double, uint32_t i) {
  uint32_t z = (uint32_t)__builtin_ctz(i), j = i >> z;
  double f = 1.0 / (double)((z + 1) * (z + 1));
  if (i <= 3) return 1.0;
  if (j == 1 || j == 3) return f;
  if (j == 9) return 0.2 * f;
  return 0.0;
}

WaveTableHost build(PhaseFn phase, AmpFn amp) {  // src/wavetable.rs:87-123
  WaveTableHost t;
  double pitch = 20.0;
  const double p_factor = pow(2.0, 1.0 / 4.0);
  float max_amplitude = 0.0f;
  while (pitch <= 20000.0) {
    std::vector<float> w = make_wave(pitch, phase, amp);
    for (float x : w) max_amplitude = fmaxf(max_amplitude, fabsf(x));
    t.pitch.push_back((float)pitch);
    t.off.push_back((int)t.data.size());
    t.len.push_back((int)w.size());
    t.data.insert(t.data.end(), w.begin(), w.end());
    pitch *= p_factor;
  }
  if (max_amplitude > 0.0f) { const float z = 1.0f / max_amplitude; for (float& x : t.data) x *= z; }
  return t;
}

}  // namespace

const WaveTableHost& global_wavetable(int kind) {
  static WaveTableHost tables[6];
  static std::once_flag once[6];
  std::call_once(once[kind], [kind]() {
    switch (kind) {
      case 0: tables[0] = build(ph_saw, am_saw); break;
      case 1: tables[1] = build(ph_zero, am_square); break;
      case 2: tables[2] = build(ph_tri, am_tri); break;
      case 3: tables[3] = build(ph_organ, am_organ); break;
      case 4: tables[4] = build(ph_organ, am_softsaw); break;
      default: tables[5] = build(ph_zero, am_hammond); break;
    }
  });
  return tables[kind];
}

// Device image of a wavetable set: each table is stored as [t[len-1]] t[0..len) [t[0] t[1]], so the four taps of the
// cubic read (i1-1, i1, i1+1, i1+2, all modulo len; src/wavetable.rs:125-155) are four CONSECUTIVE floats and the kernel
// needs one masked index per read instead of four. `off` points at t[0]; the total is padded to 16 bytes for TMA.
const WaveTableHost& device_wavetable(int kind) {
  static WaveTableHost tables[6];
  static std::once_flag once[6];
  std::call_once(once[kind], [kind]() {
    const WaveTableHost& t = global_wavetable(kind);
    WaveTableHost& d = tables[kind];
    d.pitch = t.pitch; d.len = t.len;
    for (size_t i = 0; i < t.off.size(); i++) {
      const float* w = t.data.data() + t.off[i];
      const int len = t.len[i];
      d.data.push_back(w[len - 1]);
      d.off.push_back((int)d.data.size());
      d.data.insert(d.data.end(), w, w + len);
      d.data.push_back(w[0]);
      d.data.push_back(w[1 % len]);
    }
    while (d.data.size() & 3) d.data.push_back(0.0f);
  });
  return tables[kind];
}

}  // namespace host
}  // namespace fdsp
