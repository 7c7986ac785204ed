// AOT instances: config 3a (noise >> SVF), headline (saw >> SVF), config 3b (biquad_bank fed by 8 seeded noises).
#include "../dsp/launch.cuh"
namespace fdsp { namespace host {
FDSP_INSTANCES(filter,
    FDSP_REG(Pipe<Noise, FixedSvf>),
    FDSP_REG(Pipe<SawHz, FixedSvf>),
    FDSP_REG(Pipe<Multi<30, 0, 8, Noise>, BiquadBank>))
}}
