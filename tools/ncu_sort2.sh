export H2B_MSM_PRECOMP=${SPACING:-22} H2B_MSM_SORT2_MIN_LOG=20
ncu --set full --clock-control none --import-source on -k regex:"msm_decompose_kernel|msm_partition_kernel|msm_place" --launch-skip 10 -c 5 -o gpurun_out/r02_sort2_c${SPACING:-22} python tools/ncu_target.py 24 3 > gpurun_out/ncu_sort2.log 2>&1
tail -3 gpurun_out/ncu_sort2.log
ncu -i gpurun_out/r02_sort2_c${SPACING:-22}.ncu-rep --page raw --csv > gpurun_out/r02_sort2_c${SPACING:-22}_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02_sort2_c${SPACING:-22}_raw.csv')))
hdr=rows[0]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.per_cycle_active','launch__registers_per_thread','launch__grid_size','launch__block_size','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__inst_executed.sum']
idx=[hdr.index(w) for w in want if w in hdr]
for r in rows[2:]:
    print({hdr[i].split('.')[0][-40:]: r[i][:60] for i in idx})
PY
