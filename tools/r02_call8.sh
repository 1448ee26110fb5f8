mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_c8.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest_gpu_c8.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_n1_launches.csv python tools/n1_breakdown.py > gpurun_out/r02_n1_breakdown.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02_n1_launches.csv")) if len(r)>10 and r[0].isdigit()]
half=len(rows)//2
for r in rows[half:]:
    print(f"{float(r[-1])/1e6:9.3f} ms  {r[4][:90]}")
PY
timeout 600 python tools/hbm_kernels.py 6553 > gpurun_out/r02_hbm_kernels2.jsonl 2> gpurun_out/r02_hbm_kernels2.err; echo "hbm rc=$?"; cut -c1-200 gpurun_out/r02_hbm_kernels2.jsonl
