// Small host-only entry points of the C ABI.
#include "common.cuh"
#include "../../include/mmt_b200.h"

extern "C" int mmt_abi_version(int* sm) {
  if (sm) *sm = 100;
  return 1;
}

namespace mmt { int g_pdl_enabled = 1; }

extern "C" int mmt_config_pdl(int enable) {
  const int prev = mmt::g_pdl_enabled;
  mmt::g_pdl_enabled = enable ? 1 : 0;
  return prev;
}

namespace mmt { extern int g_cluster4_enabled; extern int g_small_gemm_sms; }

extern "C" int mmt_config_small_gemm_sms(int sms) {
  const int prev = mmt::g_small_gemm_sms;
  mmt::g_small_gemm_sms = sms > 0 ? sms : 0;
  return prev;
}

extern "C" int mmt_config_cluster4(int enable) {
  const int prev = mmt::g_cluster4_enabled;
  mmt::g_cluster4_enabled = enable ? 1 : 0;
  return prev;
}
