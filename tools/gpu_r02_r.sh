cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
N=$(nvidia-smi -L | wc -l)
for mode in 1 0; do
THZ_BENCH_FUSED_REDUCE=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$mode bench.py --gpus $N --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r02/bench_fused${mode}_n$N.json 2> gpurun_out/r02/bench_fused${mode}_n$N.err
python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/r02/bench_fused${mode}_n$N.json') if l.startswith('{')][0]
print('fused=$mode', d['n_gpus'], d['ms_per_step'], d['value'], d['config']['parallelism'][:90])
PY
done
tail -3 gpurun_out/r02/bench_fused1_n$N.err
