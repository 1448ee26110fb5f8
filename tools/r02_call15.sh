mkdir -p gpurun_out
for band in 4 8 16 32 64; do
  PPB_BAND_TILES=$band timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:query_kernel -s 2 -c 1 --csv --log-file gpurun_out/r02_band_$band.csv python tools/kernel_time.py 100000 rand > /dev/null 2>&1
done
python - <<'PY'
import csv,glob
for band in (4,8,16,32,64):
    rows=[r for r in csv.reader(open(f"gpurun_out/r02_band_{band}.csv")) if len(r)>10 and r[0].isdigit()]
    print("band", band, {r[-3]: r[-1]+" "+r[-2] for r in rows})
PY
