mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 280 $TR --master-port 29561 bench.py --gpus 8 --config cfg5 --steps 2 > gpurun_out/r02_cfg5_8gpu.json 2> gpurun_out/r02_cfg5_8gpu.err; echo "cfg5 rc=$?"; grep "^{" gpurun_out/r02_cfg5_8gpu.json | cut -c1-1700; grep -v "^W\|^\[W\|\*\*\*\|OMP_NUM" gpurun_out/r02_cfg5_8gpu.err | tail -4
timeout 330 $TR --master-port 29562 bench.py --gpus 8 --config cfg4 --steps 2 > gpurun_out/r02_cfg4_8gpu.json 2> gpurun_out/r02_cfg4_8gpu.err; echo "cfg4 rc=$?"; grep "^{" gpurun_out/r02_cfg4_8gpu.json | cut -c1-2200; grep -v "^W\|^\[W\|\*\*\*\|OMP_NUM" gpurun_out/r02_cfg4_8gpu.err | tail -4
