cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
for v in "" variants/libthzdoe_k2pf222.so variants/libthzdoe_k2pf444.so ""; do
THZ_LIB=$v THZ_BENCH_LONG=0 timeout 600 python bench.py --no-cpu-baseline --no-secondary --steps 10 --warmup 3 2>gpurun_out/r02/bench_x.err | grep "^{" > gpurun_out/r02/bench_x.json
python -c "
import json; d=json.load(open('gpurun_out/r02/bench_x.json'))
print('lib=$v', d['ms_per_step'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})"
done
THZ_LIB=variants/libthzdoe_k2pf222.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "metric_shape or reference_vectors" 2>&1 | tail -1
