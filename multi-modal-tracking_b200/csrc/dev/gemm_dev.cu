// Developer-only translation unit: compiled into libmmt_b200_dev.so (build.py --dev, -DMMT_GEMM_DEV), never into the
// shipped libmmt_b200.so.  Hooks for tools/bench_gemm.py.
namespace mmt { extern long long* g_gemm_dbg; }

// Point the next mmt_gemm_bf16 launches at a device buffer of 4 int64 per CTA {total cycles of the MMA thread, cycles
// stalled on the epilogue, cycles stalled on TMA data, tiles}; nullptr switches the counters off.
extern "C" int mmt_dev_gemm_timing(long long* dbg) {
  mmt::g_gemm_dbg = dbg;
  return 0;
}
