#!/bin/bash
# Round 2: the GPU-box visits that produced profiles/r02_* (each block is one `gpurun [--gpus N] -- bash -c '...'`).
#
# --- 1 GPU -------------------------------------------------------------------------------------------------------------
#   python -m pytest tests -m gpu -x -q ; python __graft_entry__.py --smoke
#   python bench.py ; python bench.py --impl reference --steps 2 --warmup 1 ; python bench.py --config cfg2
#   nvcc -O2 -std=c++17 -o /tmp/host_floor tools/host_floor.cu -lpthread && /tmp/host_floor 4            # host-side floors
#   PPB_HOST_TRACE=1 python tools/e2e_dropin.py 100000 1                                                  # first / reuse / steady / np.empty
#   (kernel variants: at commit bfc6aad — the variant branches were deleted from the source after they were measured)
#   tools/build_variants.sh dual3:"-DPPB_DUAL_RING=1 -DPPB_STAGES=3" inc:"-DPPB_STAGE_INC=1" \
#       incprobe:"-DPPB_STAGE_INC=1 -DPPB_EARLY_PROBE=1" defer:"-DPPB_DEFER_PACK=1" all3:"-DPPB_STAGE_INC=1 -DPPB_EARLY_PROBE=1 -DPPB_DEFER_PACK=1"
#   PPB_LIB=variants/<v>.so python tools/kernel_time.py 100000 [rand]                                     # kernel variants
#   python tools/mixbench.py                                                                               # instruction-mix ceiling
#   python tools/hbm_kernels.py 6553                                                                       # N1/N2/N3 bandwidths
#   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/n1.csv python tools/n1_breakdown.py
#   ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pack_kernel|ytab_kernel|query_kernel|microbench_kernel" \
#       -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1
#   ncu --set full --clock-control none --import-source on -k regex:query_kernel -s 2 -c 1 -f -o gpurun_out/qk python tools/kernel_time.py 100000 rand
#   PPB_BAND_TILES=<b> ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum \
#       --clock-control none -k regex:query_kernel -s 2 -c 1 --csv --log-file gpurun_out/band_<b>.csv python tools/kernel_time.py 100000 rand
#       (also with PPB_DEBUG_SKIP_EPILOGUE=1, PPB_STREAM_STORES=1, PPB_A_POLICY=1, PPB_B_POLICY=2; and `python tools/perf_shapes.py cfg5`)
#   compute-sanitizer --tool memcheck|racecheck python tools/sanitize_small.py
# --- 2 GPUs ------------------------------------------------------------------------------------------------------------
#   python -m pytest tests -m gpu -x -q                      # the multi-device host-call tests need two devices
#   /tmp/host_floor 2 2 quick
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3
# --- 8 GPUs ------------------------------------------------------------------------------------------------------------
#   /tmp/host_floor 2 8 quick
#   TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
#   $TR --master-port 29551 bench.py --gpus 8 --steps 3
#   $TR --master-port 29561 bench.py --gpus 8 --config cfg5 --steps 2
#   $TR --master-port 29562 bench.py --gpus 8 --config cfg4 --steps 2
#   PPB_HOST_PIN_AFTER=1 PPB_HOST_TRACE=1 python tools/e2e_labels.py 125000 50000 8 4 1000000      # streamed labels, one process
#   (1 / 2 GPUs: PPB_HOST_TRACE=1 python tools/e2e_labels.py 125000 50000 1|2 ; PPB_HOST_CHUNK_ROWS=8000000 for ~1000 chunks per device)
echo "see the comments in this file; each line is run under gpurun"
