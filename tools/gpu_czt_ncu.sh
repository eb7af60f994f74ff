cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
THZ_CZT_IMPL=tc timeout 120 python tools/czt_accuracy.py 2048 1024 16 > gpurun_out/plain_czt.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:toeplitz_gemm_tc -c 2 -o gpurun_out/prof_r01_czt3 python tools/czt_accuracy.py 2048 1024 16 > gpurun_out/ncu_czt.log 2>&1
tail -2 gpurun_out/ncu_czt.log; cat gpurun_out/plain_czt.log | tail -3
