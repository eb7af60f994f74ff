# round 2, call B: GPU tests, bench with the secondary configs
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
rm -f gpurun_out/parity_errors.jsonl
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^PARITY" > gpurun_out/r02/pytest_gpu_b.log
tail -15 gpurun_out/r02/pytest_gpu_b.log
cp gpurun_out/parity_errors.jsonl gpurun_out/r02/parity_errors_b.jsonl
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02/bench_b.json 2> gpurun_out/r02/bench_b.err
tail -c 4000 gpurun_out/r02/bench_b.json; tail -5 gpurun_out/r02/bench_b.err
