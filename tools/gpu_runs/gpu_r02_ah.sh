mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attention_backward_ldsm -s 14 -c 1 -f -o gpurun_out/ah_attn_bwd python bench.py --mode train --steps 1 --warmup 1 > /dev/null 2>&1
ncu -i gpurun_out/ah_attn_bwd.ncu-rep --page raw --csv > gpurun_out/ah_attn_bwd_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/ah_attn_bwd_raw.csv > gpurun_out/ah_attn_bwd_summary.txt 2>&1
cat gpurun_out/ah_attn_bwd_summary.txt
python - <<'PY'
import csv
csv.field_size_limit(10**9)
rows=list(csv.reader(open('gpurun_out/ah_attn_bwd_raw.csv')))
h=rows[0]; r=rows[2]
for k in ('l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','l1tex__t_requests_pipe_lsu_mem_global_op_st.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','lts__t_sectors_op_write.sum','lts__t_sectors_op_read.sum','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','sm__inst_executed_pipe_tensor.sum','smsp__inst_executed.sum','sm__warps_active.avg.pct_of_peak_sustained_active'):
    if k in h: print(k, r[h.index(k)])
PY
