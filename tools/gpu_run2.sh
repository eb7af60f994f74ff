set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
# tuning sweeps (device-resident value only; short)
for cols in 2 4; do for chunk in 0 1 2 4; do
  echo "== K2_COLS=$cols BC_CHUNK=$chunk"; THZ_K2_COLS=$cols THZ_BC_CHUNK=$chunk timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'], {k:round(v['ms_per_step'],2) for k,v in d['roofline']['kernels'].items()})"
done; done
echo "== cached mode"; THZ_KERNEL_MODE=cached timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'], {k:round(v['ms_per_step'],2) for k,v in d['roofline']['kernels'].items()})"
# ncu: launch list, then full capture of the three pipeline kernels
python tools/profile_step.py --c 4 --steps 2 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv python tools/profile_step.py --c 4 --steps 2 > gpurun_out/ncu_list.log 2>&1
python tools/profile_step.py --c 2 --steps 1 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:thz_k -c 6 -o gpurun_out/prof_r01 python tools/profile_step.py --c 2 --steps 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
