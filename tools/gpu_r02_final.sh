# round 2 evidence run: tests + parity distances, smoke, bench (with secondary), ncu launch list + full capture of the bench step
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
rm -f gpurun_out/parity_errors.jsonl
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^PARITY" > gpurun_out/r02/final_pytest_gpu.log
tail -3 gpurun_out/r02/final_pytest_gpu.log
cp gpurun_out/parity_errors.jsonl gpurun_out/r02/final_parity_errors.jsonl
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02/final_bench_n1.json 2> gpurun_out/r02/final_bench_n1.err
tail -c 1500 gpurun_out/r02/final_bench_n1.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02/final_bench_reference.json 2>&1
python tools/profile_step.py --c 16 --steps 2 > gpurun_out/r02/final_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02/final_launches.csv python tools/profile_step.py --c 16 --steps 2 > gpurun_out/r02/final_ncu_list.log 2>&1
python tools/profile_step.py --c 16 --steps 1 > gpurun_out/r02/final_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:thz_p2_k -c 6 -o gpurun_out/r02/prof_r02_final python tools/profile_step.py --c 16 --steps 1 > gpurun_out/r02/final_ncu_full.log 2>&1
tail -2 gpurun_out/r02/final_ncu_full.log
# the report itself can exceed what gpurun brings back (64 MiB): export the two pages that are read afterwards, drop the report
ncu -i gpurun_out/r02/prof_r02_final.ncu-rep --page raw --csv > gpurun_out/r02/final_ncu_raw.csv 2>/dev/null
ncu -i gpurun_out/r02/prof_r02_final.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r02/final_ncu_source.csv.gz
ls -la gpurun_out/r02/prof_r02_final.ncu-rep gpurun_out/r02/final_ncu_raw.csv gpurun_out/r02/final_ncu_source.csv.gz
rm -f gpurun_out/r02/prof_r02_final.ncu-rep
