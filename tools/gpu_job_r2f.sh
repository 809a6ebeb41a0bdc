#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2f_build.log 2>&1
timeout 600 python tools/bf16_error_table.py > gpurun_out/r2f_bf16_errors.md 2>&1; cat gpurun_out/r2f_bf16_errors.md
timeout 600 python -m pytest tests/test_forward_gpu.py -x -q -s -k "plain_corner or template_cache" > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "HEAD_TYPE|passed|failed" gpurun_out/r2f_pytest.log
