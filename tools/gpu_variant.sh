# usage: THZ_VARIANT_FLAGS="-D..." bash tools/gpu_variant.sh   -- rebuilds on the GPU box with extra nvcc flags and runs the bench
cd $GRAFT_REPO_ROOT
for v in "$@"; do
echo "variant: $v"
THZ_NVCC_EXTRA="$v" python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
THZ_NVCC_EXTRA="$v" timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>&1 | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})"
done
