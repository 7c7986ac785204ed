// AOT instances: config 4 dry chain (saw >> moog * adsr_live >> pan), plain and stage-pipelined (dsp/bank_kernel_st.cuh).
#include "../dsp/launch.cuh"
namespace fdsp { namespace host {
FDSP_INSTANCES(sub,
    FDSP_REG_ST(SubtractiveDry))
}}
