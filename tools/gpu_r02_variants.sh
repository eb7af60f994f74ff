# A/B of environment variants on the bench step: prints ms/step and the per-kernel split for each
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-secondary --steps 10 --warmup 3 2>gpurun_out/r02/var_$name.err | grep "^{" > gpurun_out/r02/var_$name.json
  python - "$name" <<'PY'
import json,sys
name=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r02/var_%s.json'%name).read())
    print("%-28s %.3f ms/step frac %.3f e2e %.2f mode %s"%(name,d['ms_per_step'],d['roofline']['step']['frac'],d['e2e']['ms_per_step'],d['config'].get('kernel_mode_resolved')),{k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
except Exception as e:
    print(name,"FAILED",e, open('gpurun_out/r02/var_%s.err'%name).read()[-600:])
PY
}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run default
run nok2fast THZ_NO_K2FAST=1
run cached THZ_KERNEL_MODE=cached
run cached_nok2fast THZ_KERNEL_MODE=cached THZ_NO_K2FAST=1
python tools/profile_donn.py --b 1024 --events 2>&1 | tail -2
python tools/profile_donn.py --b 1024 --events --mode cached 2>&1 | tail -2
