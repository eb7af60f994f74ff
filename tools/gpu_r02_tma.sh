# row-FFT and column kernels with TMA tensor stores (thz_p2_k1t, thz_p2_k2ft): parity vs the plain-store kernels, full GPU suite, A/B bench
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
rm -f gpurun_out/parity_errors.jsonl
timeout 300 python -m pytest tests -m gpu -q -x -k "tma_store or metric_shape" 2>&1 | tail -15
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-secondary --steps 10 --warmup 3 2>gpurun_out/r02/tma_$name.err | grep "^{" > gpurun_out/r02/tma_$name.json
  python - "$name" <<'PY'
import json,sys
name=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r02/tma_%s.json'%name).read())
    print("%-12s %.3f ms/step frac %.3f e2e %.2f"%(name,d['ms_per_step'],d['roofline']['step']['frac'],d['e2e']['ms_per_step']),{k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
except Exception as e:
    print(name,"FAILED",e, open('gpurun_out/r02/tma_%s.err'%name).read()[-800:])
PY
}
run all
run all2
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
