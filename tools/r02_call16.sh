mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum --clock-control none -k regex:query_kernel -s 2 -c 1 --csv --log-file gpurun_out/r02_traffic_$name.csv python tools/kernel_time.py 100000 rand > /dev/null 2>&1; }
run default PPB_X=1
run skip_epilogue PPB_DEBUG_SKIP_EPILOGUE=1
run stream_stores PPB_STREAM_STORES=1
run a_last PPB_A_POLICY=1
run b_first PPB_B_POLICY=2
run a_last_b_first_cs PPB_A_POLICY=1 PPB_B_POLICY=2 PPB_STREAM_STORES=1
python - <<'PY'
import csv
for name in ("default","skip_epilogue","stream_stores","a_last","b_first","a_last_b_first_cs"):
    rows=[r for r in csv.reader(open(f"gpurun_out/r02_traffic_{name}.csv")) if len(r)>10 and r[0].isdigit()]
    d={r[-3]: float(r[-1].replace(",","")) for r in rows}
    print(f"{name:20s} read {d.get('dram__bytes_read.sum',0)/1e9:7.1f} GB  write {d.get('dram__bytes_write.sum',0)/1e9:6.1f} GB  {d.get('gpu__time_duration.sum',0)/1e6:7.1f} ms  L2 hit {d.get('lts__t_sector_hit_rate.pct',0):5.1f} %  L2 rd sectors {d.get('lts__t_sectors_srcunit_tex_op_read.sum',0)*32/1e9:7.0f} GB wr {d.get('lts__t_sectors_srcunit_tex_op_write.sum',0)*32/1e9:6.0f} GB")
PY
