"""raw-page CSV (exported on the GPU box with `ncu -i rep --page raw --csv`) -> compact markdown table
   python tools/ncu_raw_md.py <raw.csv> <out.md> [title]"""
import csv
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'local_load', 'smsp__inst_executed_op_local']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
out = open(sys.argv[2], 'w')
out.write('# %s\n\n(ncu --set full --clock-control none; raw page exported on the GPU box)\n\n' % (sys.argv[3] if len(sys.argv) > 3 else sys.argv[1]))
for vals in rows[2:]:
    d = dict(zip(hdr, zip(units, vals)))
    out.write('## %s\n\n| metric | value | unit |\n|---|---|---|\n' % d['Kernel Name'][1][:90])
    for k in KEYS:
        for h in d:
            if h == k or (k in ('local_load', 'smsp__inst_executed_op_local') and k in h and 'sum' in h and 'pct' not in h):
                out.write('| %s | %s | %s |\n' % (h, d[h][1], d[h][0]))
    out.write('\n')
out.close()
print(open(sys.argv[2]).read()[:6000])
