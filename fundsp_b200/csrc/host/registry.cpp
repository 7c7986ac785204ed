// Kernel registry lookup: collects the ahead-of-time instance tables (csrc/inst/*.cu).
#include "registry.h"
#include "../dsp/rt_args.h"

#include <vector>

namespace fdsp {
namespace host {

#define FDSP_DECL(name) extern const KernelEntry kInst_##name[]; extern const int kInst_##name##_n;
FDSP_DECL(osc) FDSP_DECL(filter) FDSP_DECL(net) FDSP_DECL(sub) FDSP_DECL(reverb)

namespace {

std::string expand(const std::string& s) {
  // registry keys are stringified C++ type expressions written with the typedef shorthands of dsp/launch.cuh
  static const char* names[][2] = {
      {"SubtractiveVoice", "Pipe<SubtractiveDry,Bus<MultiPass<2>,Unop<3,ReverbStereo>>>"},
      {"SubtractiveDry", "Pipe<Binop<2,Pipe<Stack<SawHz,Constant<2>>,Moog<3>>,AdsrLive>,Panner<1>>"},
      {"ReverbStereo", "Pipe<Pipe<MultiSplit<2,16>,Feedback<1,Multi<30,0,32,Pipe<Delay,Fir<3>>>>>,Binop<2,Multi<31,0,32,Panner<1>>,Constant<2>>>"},
      {"Fm", "Pipe<Unop<1,Unop<3,Unop<3,SineHz>>>,Sine>"},
      {"SineHz", "Pipe<Constant<1>,Sine>"},
      {"SawHz", "Pipe<Constant<1>,WaveSynth<0,1>>"},
  };
  std::string r;
  for (char ch : s) if (ch != ' ') r.push_back(ch);
  for (auto& n : names) {
    const std::string key = n[0];
    size_t pos = 0;
    while ((pos = r.find(key, pos)) != std::string::npos) {
      const bool lb = pos == 0 || r[pos - 1] == '<' || r[pos - 1] == ',';
      const size_t e = pos + key.size();
      const bool rb = e == r.size() || r[e] == '>' || r[e] == ',';
      if (lb && rb) r.replace(pos, key.size(), n[1]); else pos = e;
    }
  }
  return r;
}

struct Table { std::vector<const KernelEntry*> e; std::vector<std::string> key; };
const Table& table() {
  static Table t;
  if (t.e.empty()) {
    auto add = [&](const KernelEntry* a, int n) { for (int i = 0; i < n; i++) { t.e.push_back(a + i); t.key.push_back(expand(a[i].sig)); } };
    add(kInst_osc, kInst_osc_n); add(kInst_filter, kInst_filter_n); add(kInst_net, kInst_net_n); add(kInst_sub, kInst_sub_n); add(kInst_reverb, kInst_reverb_n);
  }
  return t;
}

}  // namespace

int registry_size() { return (int)table().e.size(); }
const KernelEntry* registry_at(int i) { return table().e[i]; }
const char* registry_key(int i) { return table().key[i].c_str(); }
const KernelEntry* find_kernel(const std::string& sig) {
  const Table& t = table();
  for (size_t i = 0; i < t.e.size(); i++) if (t.key[i] == sig) return t.e[i];
  return nullptr;
}


namespace {
struct AotProgram : Program {
  const KernelEntry* e;
  explicit AotProgram(const KernelEntry* e_, const std::string& key) : e(e_) {
    sig = key; IN = e->IN; OUT = e->OUT; NP = e->NP; NS = e->NS; NU = e->NU; threads = e->threads(); wave_kind = e->wave_kind();
    stages = e->launch_st ? e->stages : 1;
    has_rt = e->launch_rt != nullptr;
  }
  cudaError_t launch(const BankArgs& a, int mode, size_t table_bytes, cudaStream_t st) const override { return e->launch(a, mode, table_bytes, st); }
  cudaError_t launch_staged(const BankArgs& a, int mode, size_t table_bytes, cudaStream_t st) const override {
    return e->launch_st ? e->launch_st(a, mode, table_bytes, st) : cudaErrorInvalidValue;
  }
  cudaError_t launch_rt(const BankArgs& a, const RtArgs& rt, size_t table_bytes, cudaStream_t st) const override {
    return e->launch_rt ? e->launch_rt(a, rt, table_bytes, st) : cudaErrorInvalidValue;
  }
};
}  // namespace

std::shared_ptr<const Program> get_program(const std::string& sig, int device, std::string& err) {
  if (const KernelEntry* e = find_kernel(sig)) return std::make_shared<AotProgram>(e, sig);
  return jit_program(sig, device, err);
}

}  // namespace host
}  // namespace fdsp
