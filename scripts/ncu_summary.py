"""Print the metrics we track from an .ncu-rep (first kernel):  python scripts/ncu_summary.py file.ncu-rep [--md]"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.avg',
        'smsp__inst_executed.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__grid_size',
        'launch__registers_per_thread', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
md = '--md' in sys.argv
if md:
    print('| metric | unit | value |\n|---|---|---|')
for h, u, v in zip(hdr, units, vals):
    if h in WANT or ('pcsamp_warps_issue_stalled' in h and 'not_issued' not in h and float(v or 0) > 20000):
        print(('| %s | %s | %s |' if md else '%s %s %s') % (h, u, v))
