#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2d_build.log 2>&1
for dbg in 0 64 128 192 256 512 768; do
  echo "== MMT_GEMM_DBG=$dbg" >> gpurun_out/r2d_gemm_ln.txt
  MMT_B200_DEV_LIB=1 MMT_GEMM_DBG=$dbg timeout 300 python tools/bench_gemm_ln.py 28928 >> gpurun_out/r2d_gemm_ln.txt 2>&1
done
echo "== shipped lib" >> gpurun_out/r2d_gemm_ln.txt
timeout 300 python tools/bench_gemm_ln.py 28928 >> gpurun_out/r2d_gemm_ln.txt 2>&1
cat gpurun_out/r2d_gemm_ln.txt
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2d_pytest.log
