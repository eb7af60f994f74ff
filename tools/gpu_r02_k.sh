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
    print("%-20s %.3f ms/step frac %.3f"%(name,d['ms_per_step'],d['roofline']['step']['frac']),{k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
except Exception as e:
    print(name,"FAILED",e, open('gpurun_out/r02/var_%s.err'%name).read()[-600:])
PY
}
run base
run hints THZ_LIB=$GRAFT_REPO_ROOT/variants/libthzdoe_hints.so
run hints_st THZ_LIB=$GRAFT_REPO_ROOT/variants/libthzdoe_hints_st.so
run base2
