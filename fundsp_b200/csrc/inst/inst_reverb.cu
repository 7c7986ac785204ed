// AOT instances: reverb_stereo (32-line Hadamard FDN) and the full config 4 voice in thread-per-voice form.
// These are the generic (correctness-first) lowering; the warp-per-voice FDN kernel is csrc/dsp/fdn_kernel.cuh.
#include "../dsp/launch.cuh"
namespace fdsp { namespace host {
FDSP_INSTANCES(reverb,
    FDSP_REG(ReverbStereo),
    FDSP_REG(SubtractiveVoice))
}}
