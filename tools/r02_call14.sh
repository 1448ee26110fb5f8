mkdir -p gpurun_out
: > gpurun_out/r02_band_traffic.log
for band in 8 16 32 64 128; do
  PPB_BAND_TILES=$band timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:query_kernel -s 2 -c 1 --csv python tools/kernel_time.py 100000 rand 2>/dev/null | grep -E "query_kernel|N=100000" | awk -v b=$band -F'","' '{print "band="b, $(NF-2), $(NF-1), $NF}' >> gpurun_out/r02_band_traffic.log
  PPB_BAND_TILES=$band timeout 200 python tools/kernel_time.py 100000 rand 2>&1 | tail -1 | sed "s/^/band=$band /" >> gpurun_out/r02_band_traffic.log
done
cat gpurun_out/r02_band_traffic.log
PPB_HOST_TRACE=1 timeout 600 python tools/e2e_dropin.py 100000 1 > gpurun_out/r02_e2e_dropin_c14.jsonl 2> gpurun_out/r02_e2e_dropin_c14.err; echo "e2e rc=$?"; cut -c1-200 gpurun_out/r02_e2e_dropin_c14.jsonl; grep ppb_query_host gpurun_out/r02_e2e_dropin_c14.err | tail -4
