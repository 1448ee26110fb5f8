mkdir -p gpurun_out
timeout 300 python tools/mixbench.py > gpurun_out/r02_mixbench.jsonl 2> gpurun_out/r02_mixbench.err; echo "mixbench rc=$?"; cat gpurun_out/r02_mixbench.jsonl; tail -3 gpurun_out/r02_mixbench.err
timeout 900 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"; cat gpurun_out/r02_bench_1gpu.json; tail -15 gpurun_out/r02_bench_1gpu.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/r02_bench_ref.json
timeout 300 python bench.py --config cfg2 > gpurun_out/r02_cfg2_1gpu.json 2> gpurun_out/r02_cfg2_1gpu.err; echo "cfg2 rc=$?"; cat gpurun_out/r02_cfg2_1gpu.json; tail -3 gpurun_out/r02_cfg2_1gpu.err
timeout 300 python bench.py --config cfg4 --genomes 40000 --steps 2 > gpurun_out/r02_cfg4_small.json 2> gpurun_out/r02_cfg4_small.err; echo "cfg4 rc=$?"; cat gpurun_out/r02_cfg4_small.json; tail -3 gpurun_out/r02_cfg4_small.err
timeout 300 python bench.py --config cfg5 --genomes 6000 --steps 2 > gpurun_out/r02_cfg5_small.json 2> gpurun_out/r02_cfg5_small.err; echo "cfg5 rc=$?"; cat gpurun_out/r02_cfg5_small.json; tail -3 gpurun_out/r02_cfg5_small.err
