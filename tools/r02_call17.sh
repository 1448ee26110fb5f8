mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bucket_emit_kernel.*Iter2d|bucket_count_kernel.*Iter2d|iterate1d_classify" -c 3 -f -o gpurun_out/r02_n1_full python tools/n1_breakdown.py > gpurun_out/r02_n1_full.log 2>&1; echo "ncu n1 rc=$?"
ncu -i gpurun_out/r02_n1_full.ncu-rep --page raw --csv > gpurun_out/r02_n1_full_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/r02_n1_full_raw.csv")))
hdr=rows[0]
want=["Kernel Name","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","sm__throughput.avg.pct_of_peak_sustained_elapsed","smsp__issue_active.avg.pct","sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","launch__occupancy_limit_shared_mem","launch__occupancy_limit_registers","smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio","smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_membar_per_issue_active.ratio","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","smsp__inst_executed.sum"]
idx=[hdr.index(w) for w in want if w in hdr]
for r in rows[2:]:
    print({hdr[i]: r[i][:60] for i in idx})
PY
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:query_kernel -s 1 -c 1 --csv --log-file gpurun_out/r02_cfg5_slice_traffic.csv python tools/perf_shapes.py cfg5 > gpurun_out/r02_cfg5_slice.log 2>&1; echo "cfg5 ncu rc=$?"
grep query_kernel gpurun_out/r02_cfg5_slice_traffic.csv | awk -F'","' '{print $(NF-2), $NF}'; tail -2 gpurun_out/r02_cfg5_slice.log
