cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
rm -f gpurun_out/parity_errors.jsonl
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^PARITY" > gpurun_out/r02/pytest_gpu_f.log
tail -12 gpurun_out/r02/pytest_gpu_f.log
cp gpurun_out/parity_errors.jsonl gpurun_out/r02/parity_errors_f.jsonl
python tools/profile_donn.py --b 1024 --events 2>&1 | tail -2
timeout 300 python bench.py --no-cpu-baseline --no-secondary --steps 10 --warmup 3 2>gpurun_out/r02/bench_f.err | grep "^{" > gpurun_out/r02/bench_f.json
python -c "
import json; d=json.load(open('gpurun_out/r02/bench_f.json')); print(d['ms_per_step'], d['roofline']['step']['frac'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})"
