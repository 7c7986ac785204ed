// WAV edge of the bank: Wave::write_wav16 / write_wav32 (reference src/write.rs:24-116) byte for byte, so a
// rendered bank round-trips through the reference's file format, and a reader for the two layouts the writer produces.
#include "wavfile.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

namespace fdsp {
namespace host {

namespace {
void put32(std::vector<uint8_t>& o, uint32_t x) { o.push_back((uint8_t)x); o.push_back((uint8_t)(x >> 8)); o.push_back((uint8_t)(x >> 16)); o.push_back((uint8_t)(x >> 24)); }
void put16(std::vector<uint8_t>& o, uint16_t x) { o.push_back((uint8_t)x); o.push_back((uint8_t)(x >> 8)); }
void header(std::vector<uint8_t>& o, size_t data_length, uint16_t format, size_t channels, size_t sample_rate) {  // write_wav_header :24-52
  o.insert(o.end(), {'R', 'I', 'F', 'F'}); put32(o, (uint32_t)data_length + 36u);
  o.insert(o.end(), {'W', 'A', 'V', 'E', 'f', 'm', 't', ' '}); put32(o, 16u);
  put16(o, format); put16(o, (uint16_t)channels); put32(o, (uint32_t)sample_rate);
  const uint32_t sample_bytes = format == 1 ? 2u : 4u;
  put32(o, (uint32_t)(sample_rate * channels) * sample_bytes);
  put16(o, (uint16_t)((uint16_t)channels * (uint16_t)sample_bytes)); put16(o, (uint16_t)(sample_bytes * 8u));
  o.insert(o.end(), {'d', 'a', 't', 'a'}); put32(o, (uint32_t)data_length);
}
}  // namespace

std::string wav_encode(std::vector<uint8_t>& out, const float* planar, uint32_t channels, uint64_t length, uint64_t stride, double sample_rate, int bits) {
  if (!planar || channels == 0 || (bits != 16 && bits != 32)) return "wav: needs at least one channel and 16 or 32 bits";   // assert!(self.channels() > 0)
  if (length * channels * (uint64_t)(bits / 8) > 0xffffffd0ull) return "wav: data block over 4 GiB";
  const size_t rate = (size_t)std::llround(sample_rate);   // round(self.sample_rate()) as usize
  out.clear();
  out.reserve(44 + (size_t)length * channels * (bits / 8));
  header(out, (size_t)(bits / 8) * channels * length, bits == 16 ? 1 : 3, channels, rate);
  for (uint64_t i = 0; i < length; i++)
    for (uint32_t c = 0; c < channels; c++) {
      const float x = planar[(size_t)c * stride + i];
      if (bits == 16) {   // round(clamp11(x) * 32767.49) as i16, f32 arithmetic (:66-69)
        const float s = roundf(fminf(fmaxf(x, -1.0f), 1.0f) * 32767.49f);
        put16(out, (uint16_t)(int16_t)s);   // (a NaN sample clamps to -1 like f32::max / min do)
      } else {
        uint32_t u; memcpy(&u, &x, 4); put32(out, u);
      }
    }
  return "";
}

std::string wav_write(const char* path, const float* planar, uint32_t channels, uint64_t length, uint64_t stride, double sample_rate, int bits) {
  std::vector<uint8_t> bytes;
  std::string e = wav_encode(bytes, planar, channels, length, stride, sample_rate, bits);
  if (!e.empty()) return e;
  FILE* f = path ? fopen(path, "wb") : nullptr;
  if (!f) return std::string("wav: cannot create ") + (path ? path : "(null)");
  const bool ok = fwrite(bytes.data(), 1, bytes.size(), f) == bytes.size();
  return (fclose(f) == 0 && ok) ? "" : "wav: short write";
}

std::string wav_read(const char* path, std::vector<float>& planar, uint32_t* channels, uint64_t* length, double* sample_rate) {
  FILE* f = path ? fopen(path, "rb") : nullptr;
  if (!f) return std::string("wav: cannot open ") + (path ? path : "(null)");
  std::vector<uint8_t> b;
  uint8_t tmp[65536]; size_t n;
  while ((n = fread(tmp, 1, sizeof(tmp), f)) > 0) b.insert(b.end(), tmp, tmp + n);
  fclose(f);
  auto u16 = [&](size_t o) { return (uint32_t)b[o] | ((uint32_t)b[o + 1] << 8); };
  auto u32 = [&](size_t o) { return u16(o) | (u16(o + 2) << 16); };
  if (b.size() < 12 || memcmp(b.data(), "RIFF", 4) || memcmp(b.data() + 8, "WAVE", 4)) return "wav: not a RIFF/WAVE file";
  uint32_t fmt = 0, ch = 0, rate = 0, bits = 0; size_t data = 0, dlen = 0;
  for (size_t o = 12; o + 8 <= b.size();) {
    const uint32_t len = u32(o + 4);
    if (!memcmp(b.data() + o, "fmt ", 4) && o + 8 + 16 <= b.size()) { fmt = u16(o + 8); ch = u16(o + 10); rate = u32(o + 12); bits = u16(o + 22); }
    else if (!memcmp(b.data() + o, "data", 4)) { data = o + 8; dlen = std::min<size_t>(len, b.size() - data); break; }
    o += 8 + (size_t)len + (len & 1u);
  }
  if (!data || !ch || !((fmt == 1 && bits == 16) || (fmt == 3 && bits == 32))) return "wav: only 16-bit PCM and 32-bit float are read";
  const size_t frames = dlen / ((size_t)ch * (bits / 8));
  planar.assign((size_t)ch * frames, 0.0f);
  for (size_t i = 0; i < frames; i++)
    for (uint32_t c = 0; c < ch; c++) {
      const size_t o = data + (i * ch + c) * (bits / 8);
      if (bits == 16) planar[(size_t)c * frames + i] = (float)(int16_t)u16(o) / 32768.0f;   // the decoder convention the reference's reader (symphonia) uses
      else { uint32_t u = u32(o); float x; memcpy(&x, &u, 4); planar[(size_t)c * frames + i] = x; }
    }
  *channels = ch; *length = frames; *sample_rate = (double)rate;
  return "";
}

}  // namespace host
}  // namespace fdsp
