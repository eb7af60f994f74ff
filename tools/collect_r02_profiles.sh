# copies the outputs of tools/gpu_r02_final.sh (gpurun_out/r02/final_*) into profiles/r02_* and regenerates the ncu summaries
set -e
cd "$(dirname "$0")/.."
S=gpurun_out/r02
grep "^{" $S/final_bench_n1.json > profiles/r02_bench_n1.json
grep "^{" $S/final_bench_reference.json > profiles/r02_bench_reference.json
cp $S/final_launches.csv profiles/r02_final_launches.csv
cp $S/final_parity_errors.jsonl profiles/r02_parity_errors.jsonl
cp $S/final_pytest_gpu.log profiles/r02_pytest_gpu.log
cp $S/final_ncu_raw.csv /tmp/r02_final_raw.csv
python tools/ncu_summary.py /tmp/r02_final_raw.csv > profiles/r02_final_ncu_summary.txt
gunzip -c $S/final_ncu_source.csv.gz > /tmp/r02_final_src.csv
python tools/sass_mix.py /tmp/r02_final_src.csv thz_p2_k2f > profiles/r02_final_k2f_sass_mix.txt 2>/dev/null || true
(python tools/sass_mix.py /tmp/r02_final_src.csv thz_p2_k1; python tools/sass_mix.py /tmp/r02_final_src.csv thz_p2_k3) > profiles/r02_final_k1_k3_sass_mix.txt 2>/dev/null || true
python - <<'PY'
import csv, json
rows = list(csv.reader(open('/tmp/r02_final_raw.csv')))
h = {n: i for i, n in enumerate(rows[0])}
for r in rows[2:]:
    if 'thz_p2_k2f' in r[h['Kernel Name']]:
        def gb(name):
            v, u = float(r[h[name]]), rows[1][h[name]]
            return v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}[u]
        json.dump({"kernel": r[h["Kernel Name"]].replace("void ", ""), "fields_per_launch": 16, "dram_bytes_read": gb('dram__bytes_read.sum'),
                   "dram_bytes_write": gb('dram__bytes_write.sum'),
                   "source": "ncu --set full --clock-control none --import-source on, tools/profile_step.py --c 16 (tools/gpu_r02_final.sh; raw page exported on the GPU box)"},
                  open('profiles/r02_k2_dram_traffic.json', 'w'), indent=1)
        break
PY
wc -c profiles/r02_*
