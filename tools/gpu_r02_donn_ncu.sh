cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
python tools/profile_donn.py --b 256 --steps 1 > gpurun_out/r02/plain_donn.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:thz_p2_k --launch-skip 6 --launch-count 6 -o gpurun_out/r02/prof_r02_donn python tools/profile_donn.py --b 256 --steps 1 > gpurun_out/r02/ncu_donn.log 2>&1
tail -2 gpurun_out/r02/ncu_donn.log
