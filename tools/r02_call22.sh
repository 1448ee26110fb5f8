mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29571 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02_bench_8gpu_final.json 2> gpurun_out/r02_bench_8gpu_final.err; echo "bench8 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02_bench_8gpu_final.json") if l.startswith("{")][0])
    print("value",d["value"],"ms",d["ms_per_step"],"parity_ok",d["parity_ok"],"nccl",d["nccl_allgather"]["ms_per_step"])
    e=d["e2e"]; print({k:v for k,v in e.items() if k not in("api","first_call_note")}); print(d["d2h_floor"])
except Exception as ex: print("ERR",ex)
PY
grep "failed\|Error\|error" gpurun_out/r02_bench_8gpu_final.err | head -5
timeout 280 $TR --master-port 29572 bench.py --gpus 8 --config cfg5 --steps 2 > gpurun_out/r02_cfg5_8gpu_final.json 2> gpurun_out/r02_cfg5_8gpu_final.err; echo "cfg5 rc=$?"; grep "^{" gpurun_out/r02_cfg5_8gpu_final.json | cut -c1-1700
timeout 330 $TR --master-port 29573 bench.py --gpus 8 --config cfg4 --steps 2 > gpurun_out/r02_cfg4_8gpu_final.json 2> gpurun_out/r02_cfg4_8gpu_final.err; echo "cfg4 rc=$?"; grep "^{" gpurun_out/r02_cfg4_8gpu_final.json | cut -c1-2300; grep -v "^W\|^\[W\|\*\*\*\|OMP_NUM" gpurun_out/r02_cfg4_8gpu_final.err | tail -3
