# after the TMA kernels: re-check the tuning knobs that were set while the column kernel was bound by the L1 data pipe
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-secondary --steps 10 --warmup 3 2>gpurun_out/r02/sw_$name.err | grep "^{" > gpurun_out/r02/sw_$name.json
  python - "$name" <<'PY'
import json,sys
name=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r02/sw_%s.json'%name).read())
    print("%-12s %.3f ms/step frac %.3f"%(name,d['ms_per_step'],d['roofline']['step']['frac']),{k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})
except Exception as e:
    print(name,"FAILED",e, open('gpurun_out/r02/sw_%s.err'%name).read()[-800:])
PY
}
run base
run k2fu2 THZ_LIB=variants/libthzdoe_k2fu2.so
run k2pf50 THZ_K2_PF=50
run k2pf200 THZ_K2_PF=200
run k3pf50 THZ_K3_PF=50
run k3pf200 THZ_K3_PF=200
run base2
