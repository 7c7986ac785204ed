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
  cudaError_t (*launch)(const BankArgs& a, int mode, cudaStream_t stream);
  int (*threads)();
};

const KernelEntry* find_kernel(const std::string& sig);
int registry_size();
const KernelEntry* registry_at(int i);
cudaError_t launch_mix_reduce(const float* partial, uint32_t nparts, uint32_t outs, uint32_t n, float* mix, uint32_t mix_stride,
                              uint32_t mix_offset, int accumulate, cudaStream_t stream);

}  // namespace host
}  // namespace fdsp
