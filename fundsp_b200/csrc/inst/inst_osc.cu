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

cudaError_t launch_tree_mix(const float* rows, uint32_t V, uint32_t outs, uint32_t row_stride, uint32_t row_offset, uint32_t n, float* mix,
                            uint32_t mix_stride, uint32_t mix_offset, int pairwise, cudaStream_t st, float* scratch) {
  const unsigned total = outs * n;
  constexpr int LB = 8;                                   // 256-voice blocks (tree_mix_scratch_floats)
  if (pairwise && scratch && V >= (4u << LB)) {           // big balanced trees: block subtrees in parallel, then the top of the tree
    const uint32_t nfull = V >> LB;
    tree_mix_block_kernel<LB><<<dim3((total + 127) / 128, nfull), 128, 0, st>>>(rows, outs, row_stride, row_offset, n, scratch);
    tree_mix_kernel<<<(total + 127) / 128, 128, 0, st>>>(rows, V, outs, row_stride, row_offset, n, mix, mix_stride, mix_offset, 1, scratch, nfull, LB);
    return cudaGetLastError();
  }
  tree_mix_kernel<<<(total + 127) / 128, 128, 0, st>>>(rows, V, outs, row_stride, row_offset, n, mix, mix_stride, mix_offset, pairwise);
  return cudaGetLastError();
}
size_t tree_mix_scratch_floats(uint32_t V, uint32_t outs, uint32_t n) { return V >= (4u << 8) ? (size_t)(V >> 8) * outs * n : 0; }
}}
