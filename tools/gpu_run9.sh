cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 180 python -m pytest tests -m gpu -q -x -k "czt" > gpurun_out/pytest_czt.log 2>&1; echo "czt pytest rc=$?"; tail -25 gpurun_out/pytest_czt.log
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
