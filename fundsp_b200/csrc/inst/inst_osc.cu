// AOT instances: oscillators, config 1 (plumbing), config 2 (FM bank); also hosts the mix-down finisher.
#include "../dsp/launch.cuh"
namespace fdsp { namespace host {
FDSP_INSTANCES(osc,
    FDSP_REG(Pipe<SineHz, FixedSvf>),
    FDSP_REG(SineHz),
    FDSP_REG(SawHz),
    FDSP_REG(Noise),
    FDSP_REG(Fm))

cudaError_t launch_mix_reduce(const float* partial, uint32_t nparts, uint32_t outs, uint32_t n, float* mix, uint32_t mix_stride,
                              uint32_t mix_offset, int accumulate, cudaStream_t st) {
  const unsigned total = outs * n;
  mix_reduce_kernel<<<(total + 255) / 256, 256, 0, st>>>(partial, nparts, outs, n, mix, mix_stride, mix_offset, accumulate);
  return cudaGetLastError();
}
}}
