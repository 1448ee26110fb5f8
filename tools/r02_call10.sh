mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_c10.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_pytest_gpu_c10.log
timeout 600 python tools/hbm_kernels.py 6553 > gpurun_out/r02_hbm_kernels4.jsonl 2> gpurun_out/r02_hbm_kernels4.err; echo "hbm rc=$?"; cut -c1-170 gpurun_out/r02_hbm_kernels4.jsonl; tail -3 gpurun_out/r02_hbm_kernels4.err
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_2gpu_b.json 2> gpurun_out/r02_bench_2gpu_b.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_2gpu_b.json") if l.startswith("{")][0])
print("value",d["value"],"ms",d["ms_per_step"],"parity_ok",d["parity_ok"])
e=d["e2e"]; print({k:v for k,v in e.items() if k not in("api","first_call_note")}); print(d["d2h_floor"])
PY
grep "step times\|parity\|failed" gpurun_out/r02_bench_2gpu_b.err | head
timeout 600 $TR --master-port 29542 bench.py --gpus 2 --config cfg4 --genomes 40000 --steps 2 > gpurun_out/r02_cfg4_small_2gpu.json 2> gpurun_out/r02_cfg4_small_2gpu.err; echo "cfg4 rc=$?"; grep "^{" gpurun_out/r02_cfg4_small_2gpu.json | cut -c1-1800; grep -v "^W\|^\[W\|\*\*\*\|OMP_NUM" gpurun_out/r02_cfg4_small_2gpu.err | tail -5
timeout 600 $TR --master-port 29543 bench.py --gpus 2 --config cfg5 --genomes 6000 --steps 2 > gpurun_out/r02_cfg5_small_2gpu.json 2> gpurun_out/r02_cfg5_small_2gpu.err; echo "cfg5 rc=$?"; grep "^{" gpurun_out/r02_cfg5_small_2gpu.json | cut -c1-1500; grep -v "^W\|^\[W\|\*\*\*\|OMP_NUM" gpurun_out/r02_cfg5_small_2gpu.err | tail -5
