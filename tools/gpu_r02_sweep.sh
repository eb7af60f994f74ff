# A/B of experiment libraries (variants/) against the default library; the default run doubles as the check of the in-tree build
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
for v in variants/libthzdoe_*.so; do n=$(basename $v .so); n=${n#libthzdoe_}; run $n THZ_LIB=$v; THZ_LIB=$v timeout 200 python -m pytest tests -m gpu -q -x -k "tma_store" 2>&1 | tail -1; done
run base2
