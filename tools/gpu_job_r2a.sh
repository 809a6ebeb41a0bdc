#!/bin/bash
# round 2, GPU job A: new parity tests, the bench with the eager-GPU / reference baselines, cuBLAS kernel probe
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2a_build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2a_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.err; echo "ref rc=$?"
timeout 300 python tools/cublas_probe.py > gpurun_out/r2a_cublas.txt 2>&1; echo "probe rc=$?"
cat gpurun_out/r2a_cublas.txt | cut -c1-600
