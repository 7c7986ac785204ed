// Plain-data control block of the resident process() kernel (dsp/bank_kernel_rt.cuh), shared with the host runtime.
#pragma once
#ifndef __CUDACC_RTC__
#include <cstdint>
#endif

namespace fdsp {

constexpr uint32_t RT_QUIT = 0xffffffffu, RT_EXITED = 0xfffffffeu;
constexpr uint32_t RT_POLL_LIMIT_HOST = 1u << 20;     // ~1 s of doorbell polls over PCIe, then the kernel leaves by itself
constexpr uint32_t RT_POLL_LIMIT_RELAY = 1u << 24;    // safety net of the CTAs that poll the relay word (longer than the leader's)

struct RtCtl {                      // mapped pinned host memory, one page
  volatile uint32_t doorbell;       // host -> device: sequence number of the request (RT_QUIT: leave)
  volatile uint32_t size;           // samples of the request (1..64)
  uint32_t pad0[14];
  volatile uint32_t done;           // device -> host: sequence number of the last finished request, or RT_EXITED
  uint32_t pad1[15];
  float in[8][64];                  // shared bank inputs of the block
  float out[8][64];                 // mix of the block
};
struct RtArgs {
  RtCtl* ctl;                       // device view of the mapped page
  uint32_t* relay;                  // device memory: [0] sequence number, [1] size, [2] exit counter
  uint32_t first_seq;               // the kernel starts having "seen" first_seq - 1
};

}  // namespace fdsp
