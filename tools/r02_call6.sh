mkdir -p gpurun_out
PPB_BENCH_NO_SAMPLER=1 timeout 600 python bench.py --steps 4 > gpurun_out/r02_bench_nosampler.json 2> gpurun_out/r02_bench_nosampler.err; echo "rc=$?"; grep "step times\|parity" gpurun_out/r02_bench_nosampler.err
timeout 600 python bench.py --steps 4 > gpurun_out/r02_bench_sampler.json 2> gpurun_out/r02_bench_sampler.err; echo "rc=$?"; grep "step times\|parity" gpurun_out/r02_bench_sampler.err
python - <<'PY'
import json
for f in ("nosampler","sampler"):
    d=json.loads([l for l in open(f"gpurun_out/r02_bench_{f}.json") if l.startswith("{")][0])
    print(f, d["ms_per_step"], d["no_table"]["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["ms_per_step"], d["e2e"]["first_call_ms"], d["clocks"])
PY
