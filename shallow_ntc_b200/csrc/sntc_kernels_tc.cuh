// tcgen05 / TMEM / TMA band-GEMM path (SNTC_PRECISION_TC_F16X3).  Placeholder until the kernels land:
// requesting the precision fails loudly at finalize time instead of silently running fp32.
#pragma once
#include <string>
#include <vector>
#include "sntc_plan.hpp"

namespace sntc {

enum { TC_OK = 0, TC_NOT_HANDLED = 1, TC_ERROR = 2 };

struct TcDriver {
  void init() {}
};

struct TcModelState {
  void release() {}
};

inline bool tc_finalize(TcDriver&, TcModelState&, Transform*, Transform*, const HostWeights&, std::vector<void*>&, std::string* err) {
  *err = "tensor-core path not built";
  return false;
}

inline int tc_run_transform(TcDriver&, TcModelState&, Transform&, bool, const float*, int, int, int, float*, uint8_t*, float*, int, int,
                            cudaStream_t, uint64_t*, std::string*) {
  return TC_NOT_HANDLED;
}

inline int tc_run_hyper_fused(TcDriver&, TcModelState&, Transform&, const float*, int, int, int, const void*, int, float*, uint8_t*,
                              float, bool, cudaStream_t, uint64_t*, std::string*) {
  return TC_NOT_HANDLED;
}

}  // namespace sntc
