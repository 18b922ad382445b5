"""Summarise an ncu report for the UMMA conv kernel: key raw metrics + stall samples per code region.
  python scripts/ncu_summary.py gpurun_out/prof.ncu-rep"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread',
        'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__t_sectors_op_read.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__cycles_elapsed.max', 'sm__cycles_active.avg',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct',
        'smsp__cycles_active.avg', 'sm__warps_active.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    print('---')
    for i, h in enumerate(hdr):
        if any(h.endswith(w) or h == w for w in want):
            print(f'{h} [{units[i]}] = {r[i]}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if len(r) < len(hdr):
        if r and r[0] == 'Kernel Name':
            break
        continue
    data.append(r)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
ts = sum(int(r[ci['# Samples']]) for r in data)
print('total samples', ts, 'instructions', len(data))
B = int(sys.argv[2]) if len(sys.argv) > 2 else 100
for b in range(0, len(data), B):
    ch = data[b:b + B]
    s = sum(int(r[ci['# Samples']]) for r in ch)
    ex = sum(int(r[ci['Instructions Executed']]) for r in ch)
    ops = collections.Counter(re.sub(r'^@!?U?P\d+\s+', '', r[ci['Source']].strip()).split()[0].split('.')[0] for r in ch)
    note = {o: ops[o] for o in ('LDTM', 'UTCHMMA', 'UBLKCP', 'REDG', 'LDG', 'FFMA', 'STS', 'LDS', 'SYNCS', 'F2FP', 'UTCBAR', 'STTM') if ops[o]}
    st = collections.Counter()
    for r in ch:
        for h in stalls:
            st[h] += int(r[ci[h]])
    print(b, f'{100 * s / ts:5.1f}%', ex, note, st.most_common(3))
top = sorted(enumerate(data), key=lambda x: -int(x[1][ci['# Samples']]))[:25]
for idx, r in sorted(top):
    s = sorted(((h, int(r[ci[h]])) for h in stalls if int(r[ci[h]]) > 0), key=lambda x: -x[1])[:2]
    print(idx, r[ci['Source']].strip()[:64].ljust(64), r[ci['# Samples']], r[ci['Instructions Executed']], s)
