cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
rm -f gpurun_out/parity_errors.jsonl
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for lut in 0 1; do
THZ_NO_DOE_LUT=$((1-lut)) THZ_BENCH_LONG=0 timeout 600 python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>gpurun_out/r02/bench_lut$lut.err | grep "^{" > gpurun_out/r02/bench_lut$lut.json
python -c "
import json; d=json.load(open('gpurun_out/r02/bench_lut$lut.json')); s=d['secondary']
print('lut=$lut', d['ms_per_step'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()}, 'C2', s['c2_step_1000_to_2000_8level']['eager_ms'], s['c2_step_1000_to_2000_8level']['cuda_graph_ms'], 'DONN', s['c4_donn_3layer_200_batch1024']['ms_per_step'])"
done
python examples/four_focal_spots.py --graph --iters 200 2>&1 | tail -1
