# plain-output instantiation of the row-iFFT kernels (MODE 3): full GPU suite, then A/B of the bench incl. the secondary configs
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
rm -f gpurun_out/parity_errors.jsonl
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
run() {
  name=$1; shift
  env "$@" THZ_BENCH_LONG=0 timeout 600 python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>gpurun_out/r02/k3p_$name.err | grep "^{" > gpurun_out/r02/k3p_$name.json
  python -c "
import json; d=json.load(open('gpurun_out/r02/k3p_$name.json')); s=d['secondary']
print('$name', round(d['ms_per_step'],4), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()}, 'C2', round(s['c2_step_1000_to_2000_8level']['cuda_graph_ms'],4), 'DONN', round(s['c4_donn_3layer_200_batch1024']['ms_per_step'],3), 'C5', round(s['c5_single_gpu_16384_padded']['fwd_bwd_ms'],3))"
}
run plain3
run mode0 THZ_NO_K3PLAIN=1
