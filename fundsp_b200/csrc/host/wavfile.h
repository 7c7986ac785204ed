// WAV files in the reference's layout (src/write.rs): 44-byte header, interleaved 16-bit PCM or 32-bit float.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace fdsp {
namespace host {

// planar[c * stride + i]; bits 16 (Wave::write_wav16) or 32 (Wave::write_wav32). Returns "" or an error text.
std::string wav_encode(std::vector<uint8_t>& out, const float* planar, uint32_t channels, uint64_t length, uint64_t stride, double sample_rate, int bits);
std::string wav_write(const char* path, const float* planar, uint32_t channels, uint64_t length, uint64_t stride, double sample_rate, int bits);
std::string wav_read(const char* path, std::vector<float>& planar, uint32_t* channels, uint64_t* length, double* sample_rate);

}  // namespace host
}  // namespace fdsp
