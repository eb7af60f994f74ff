cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
rm -f gpurun_out/parity_errors.jsonl
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python tools/profile_donn.py --b 1024 --events 2>&1 | tail -2
timeout 300 python bench.py --no-cpu-baseline --no-secondary --steps 10 --warmup 3 2>gpurun_out/r02/bench_j.err | grep "^{" > gpurun_out/r02/bench_j.json
python -c "
import json; d=json.load(open('gpurun_out/r02/bench_j.json')); print(d['ms_per_step'], d['roofline']['step']['frac'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})"
