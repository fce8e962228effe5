"""Print the handful of ncu raw metrics that matter for the rollout kernel. usage: ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct', 'smsp__average_warps_issue_stalled', 'sass__inst_executed_local', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct',
        'sm__inst_executed_pipe_lsu', 'sm__inst_executed_pipe_alu', 'sm__inst_executed_pipe_fma', 'sm__inst_executed_pipe_xu', 'sm__inst_executed_pipe_uniform',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed_op_shared', 'sm__cycles_active.avg', 'smsp__pcsamp_sample_buffer',
        'sass__thread_inst_executed_true_per_opcode', 'sm__sass_thread_inst_executed_op_dfma_pred_on.sum', 'sm__sass_thread_inst_executed_op_dmul_pred_on.sum', 'sm__sass_thread_inst_executed_op_dadd_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_fp64_pred_on.sum', 'sm__sass_thread_inst_executed_ops_dadd_dmul_dfma_pred_on.sum']
for ri in range(2, len(rows)):
    print('--- launch', ri - 2, rows[ri][hdr.index('Kernel Name')][:60] if 'Kernel Name' in hdr else '')
    for h, u, v in zip(hdr, rows[1], rows[ri]):
        if any(h.startswith(w) for w in want):
            if 'stalled' in h:
                try:
                    if float(v) < 0.05: continue
                except ValueError:
                    pass
            if h.endswith(('.max.pct_of_peak_sustained_active', '.min.pct_of_peak_sustained_active', 'per_second', '.pct_of_peak_sustained_elapsed')) and 'fp64' not in h: continue
            print('  ', h, u, v)
