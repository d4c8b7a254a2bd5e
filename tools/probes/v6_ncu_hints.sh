# DRAM traffic / instruction count of K1 v6 under the L2 eviction hints (bit 0: old match sources evict_first,
# bit 1: output stores evict_last, bit 2: compressed input evict_first)
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio
python tools/k1_probe.py --size-mib 4096 --tunings 60 --reps 1 > gpurun_out/plain.log 2>&1 || exit 1
for h in 0 1 3 5; do   # (needs a build with -DLZ4B200_V6_HINTS=$h each; the runtime switch was removed after this measurement)
  LZ4B200_V6_HINTS=$h ncu --metrics $M --clock-control none -k regex:decode_blocks_v6 -s 1 -c 1 --csv --log-file gpurun_out/v6_hints_$h.csv python tools/k1_probe.py --size-mib 4096 --tunings 60 --reps 1 > /dev/null 2>&1
done
