#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2o_build.log 2>&1
for i in 1 2; do
echo "== dev lib: 5 ring stages (8 KB epilogue region per warp); only the plain columns are meaningful" 
MMT_B200_DEV_LIB=1 timeout 300 python tools/bench_gemm_ln.py 28928 2>&1 | grep -E "^(qkv|fc1|proj|fc2)" | sed -e "s/'ln_[a-z]*': \[[^]]*\], //"
echo "== shipped: 4 ring stages"
timeout 300 python tools/bench_gemm_ln.py 28928 2>&1 | grep -E "^(qkv|fc1|proj|fc2)"
done 2>&1 | tee gpurun_out/r2o_stages.txt
