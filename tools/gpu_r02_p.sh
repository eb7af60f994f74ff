cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
rm -f gpurun_out/parity_errors.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "long or split or any" 2>&1 | tail -30
cat gpurun_out/parity_errors.jsonl | grep -i "long\|split"
