// Small host-only entry points of the C ABI.
#include "common.cuh"
#include "../../include/mmt_b200.h"

extern "C" int mmt_abi_version(int* sm) {
  if (sm) *sm = 100;
  return 1;
}
