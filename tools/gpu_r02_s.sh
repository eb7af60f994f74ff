cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
for v in 2 3 4; do
THZ_T1_LOG2=$v timeout 300 python bench.py --no-cpu-baseline --no-secondary --steps 10 --warmup 3 2>/dev/null | grep "^{" > gpurun_out/r02/bench_t1_$v.json
python -c "
import json; d=json.load(open('gpurun_out/r02/bench_t1_$v.json')); print('T1_LOG2=$v', d['ms_per_step'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})"
done
