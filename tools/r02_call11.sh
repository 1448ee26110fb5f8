mkdir -p gpurun_out
(nvidia-smi topo -m; nproc; free -g; lscpu | grep -E "Model name|Socket|NUMA") > gpurun_out/r02_topo_8gpu.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 330 $TR --master-port 29551 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err; echo "bench8 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02_bench_8gpu.json") if l.startswith("{")][0])
    print("value",d["value"],"ms",d["ms_per_step"],"parity_ok",d["parity_ok"],"nccl",d["nccl_allgather"]["ms_per_step"])
    e=d["e2e"]; print({k:v for k,v in e.items() if k not in("api","first_call_note")}); print(d["d2h_floor"]); print(d["rectangular"])
except Exception as ex: print("ERR",ex)
PY
grep "step times\|parity\|failed\|Error\|error" gpurun_out/r02_bench_8gpu.err | head -12
nvcc -O2 -std=c++17 -o /tmp/host_floor tools/host_floor.cu -lpthread 2>/dev/null && timeout 150 /tmp/host_floor 2 8 quick > gpurun_out/r02_host_floor_8gpu.jsonl 2> gpurun_out/r02_host_floor_8gpu.err; echo "floor rc=$?"; cat gpurun_out/r02_host_floor_8gpu.jsonl
