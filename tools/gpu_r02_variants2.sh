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
run fast_u1
run nok2fast THZ_NO_K2FAST=1
run t1log1 THZ_NO_K2FAST=1 THZ_T1_LOG2=1
run cached_u1 THZ_KERNEL_MODE=cached
python tools/profile_step.py --c 4 --steps 1 > gpurun_out/r02/plain_prof.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:thz_p2_k2 -c 2 -o gpurun_out/r02/prof_r02_k2f python tools/profile_step.py --c 4 --steps 1 > gpurun_out/r02/ncu_k2f.log 2>&1
tail -2 gpurun_out/r02/ncu_k2f.log
