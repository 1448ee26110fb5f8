mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_c21.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_gpu_c19.log
timeout 600 python tools/hbm_kernels.py 6553 > gpurun_out/r02_hbm_kernels8.jsonl 2> gpurun_out/r02_hbm_kernels8.err; echo "hbm rc=$?"; cut -c1-170 gpurun_out/r02_hbm_kernels8.jsonl; tail -3 gpurun_out/r02_hbm_kernels8.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_n1_launches5.csv python tools/n1_breakdown.py > gpurun_out/r02_n1_breakdown5.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02_n1_launches5.csv")) if len(r)>10 and r[0].isdigit()]
half=len(rows)//2
for r in rows[half:]:
    print(f"{float(r[-1])/1e6:9.3f} ms  {r[4][:80]}")
PY
