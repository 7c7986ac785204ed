// AOT instances: the four voice classes of config 5 (dynamic Net), each `>> pan(p)`.
#include "../dsp/launch.cuh"
namespace fdsp { namespace host {
FDSP_INSTANCES(net,
    FDSP_REG(Pipe<Pipe<SineHz, FixedSvf>, Panner<1>>),
    FDSP_REG_ST(Pipe<Pipe<SawHz, Moog<1>>, Panner<1>>),
    FDSP_REG(Pipe<Pipe<Noise, FixedSvf>, Panner<1>>),
    FDSP_REG(Pipe<Pipe<Fm, FixedSvf>, Panner<1>>))
}}
