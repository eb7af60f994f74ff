# round 2, call A: full GPU test suite (no -x: see every failure), measured parity distances, bench line
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
rm -f gpurun_out/parity_errors.jsonl
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^PARITY" > gpurun_out/r02/pytest_gpu_a.log
tail -40 gpurun_out/r02/pytest_gpu_a.log
cp gpurun_out/parity_errors.jsonl gpurun_out/r02/parity_errors_a.jsonl
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02/bench_a.json 2> gpurun_out/r02/bench_a.err
tail -c 3000 gpurun_out/r02/bench_a.json
