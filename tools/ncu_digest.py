"""Digest an .ncu-rep: key raw metrics, stall-reason totals, hottest SASS lines.
    python tools/ncu_digest.py gpurun_out/prof.ncu-rep [n_top]"""
import csv
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.max', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'lts__t_sector_hit_rate.pct',
        'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'sm__inst_executed_pipe_alu.sum',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_fmaheavy.sum',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum', 'smsp__cycles_active.avg']
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
for h, u, v in zip(rows[0], rows[1], rows[-1]):
    if h in KEEP:
        print(f"{h:70s} {v} {u}")
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
num = lambda r, k: int(r[ix[k]] or 0)
tot = sum(num(r, '# Samples') for r in data)
inst = sum(num(r, 'Instructions Executed') for r in data)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: sum(num(r, s) for r in data) for s in stalls}
print('samples', tot, 'warp-inst', inst)
print(sorted(agg.items(), key=lambda x: -x[1])[:9])
ops = Counter()
for r in data:
    src = [t for t in r[ix['Source']].split() if not t.startswith('@')]
    ops[src[0].split('.')[0] if src else '?'] += num(r, 'Instructions Executed')
print([(o, round(100 * n / inst, 1)) for o, n in ops.most_common(16)])
for r in sorted(data, key=lambda r: -num(r, '# Samples'))[:ntop]:
    st = sorted(((s, num(r, s)) for s in stalls), key=lambda x: -x[1])[:2]
    print(r[ix['Address']][-5:], str(num(r, '# Samples')).rjust(6), str(num(r, 'Instructions Executed')).rjust(9),
          r[ix['Source']][:64].ljust(64), st)
