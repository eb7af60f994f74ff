cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
python tools/profile_step.py --c 16 --steps 2 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01_final.csv python tools/profile_step.py --c 16 --steps 2 > gpurun_out/ncu_list.log 2>&1
python tools/profile_step.py --c 16 --steps 1 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:thz_p2_k -c 6 -o gpurun_out/prof_r01_final python tools/profile_step.py --c 16 --steps 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
