"""Key metrics of each kernel in an .ncu-rep (run where ncu is installed): python scripts/ncu_brief.py file.ncu-rep"""
import csv, subprocess, sys, io
KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__cycles_active.avg']
for f in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', f, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(h, r)); u = dict(zip(h, units))
        print('==', f, d.get('Kernel Name', '')[:90])
        for k in KEYS:
            if k in d: print('   %-70s %s %s' % (k, d[k], u[k]))
        st = sorted(((float(d[k]), k) for k in h if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and d[k]), reverse=True)[:6]
        for v, k in st: print('   stall %-64s %.2f' % (k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))
