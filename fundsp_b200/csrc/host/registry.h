// Kernel registry: maps a device program type expression (graph.h `sig`) to its launcher.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "../dsp/bank_args.h"

namespace fdsp {
namespace host {

struct KernelEntry {
  const char* sig;
  int IN, OUT, NP, NS, NU;
  // mode: FDSP_OUT_VOICES | FDSP_OUT_MIX bits
  // table_bytes > 0: stage that many bytes of wavetable data in shared memory (long renders of wavetable graphs)
  cudaError_t (*launch)(const BankArgs& a, int mode, size_t table_bytes, cudaStream_t stream);
  int (*threads)();
  int (*wave_kind)();  // first wavetable kind the program reads, -1 if none
};

const KernelEntry* find_kernel(const std::string& sig);
int registry_size();
const KernelEntry* registry_at(int i);
cudaError_t launch_mix_reduce(const float* partial, uint32_t nparts, uint32_t outs, uint32_t n, float* mix, uint32_t mix_stride,
                              uint32_t mix_offset, int accumulate, cudaStream_t stream);

}  // namespace host
}  // namespace fdsp
