#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2c_build.log 2>&1
timeout 300 python tools/bench_gemm_ln.py 28928 > gpurun_out/r2c_gemm_ln.txt 2>&1; echo "gemm_ln rc=$?"; cat gpurun_out/r2c_gemm_ln.txt
timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_native_ops_module_gpu.py tests/test_frames_gpu.py -x -q -s -k "template_cache or native_ops or framestep or run_sequences or full_size_batch_against or bf16_mode" > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "forced-keep|CE:|bs=|native_ops|passed|failed|Error" gpurun_out/r2c_pytest.log | tail -30
