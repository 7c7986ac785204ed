// AOT instances: config 4 dry chain (saw >> moog * adsr_live >> pan).
#include "../dsp/launch.cuh"
namespace fdsp { namespace host {
FDSP_INSTANCES(sub,
    FDSP_REG(SubtractiveDry))
}}
