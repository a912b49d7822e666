"""Summarise an .ncu-rep (raw + source pages) into text: python tools/ncu_summary.py rep [topN]"""
import csv, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr, units, vals = rows[0], rows[1], rows[2]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'sm__inst_executed.sum.per_cycle_elapsed', 'sm__inst_executed.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.avg.per_second',
        'launch__grid_size', 'launch__block_size', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'lts__t_bytes.sum', 'sm__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__issue_inst0.avg.pct_of_peak_sustained_active']
for h, u, v in zip(hdr, units, vals):
    if h in keys or any(h.endswith(k) for k in keys): print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines())); hdr = rows[1]; data = rows[2:]; ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot = sum(f(r, '# Samples') for r in data); ninst = sum(f(r, 'Instructions Executed') for r in data)
print(f"--- source page: {len(data)} SASS instr, {tot:.0f} samples, {ninst:.3e} warp-instr executed")
agg = defaultdict(lambda: [0, 0])
for r in data:
    t = r[ix['Source']].split(); op = t[1] if t and t[0].startswith('@') else (t[0] if t else '?')
    agg[op][0] += f(r, '# Samples'); agg[op][1] += f(r, 'Instructions Executed')
print("--- by opcode: samples% / warp-instr executed")
for op, (s, n) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    print(f"{op:30s} {100 * s / max(tot, 1):5.1f}%  {n:14.0f}")
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
print("--- stall reasons (all samples)")
for s in sorted(stalls, key=lambda s: -sum(f(r, s) for r in data))[:9]: print(f"{s:28s}{sum(f(r, s) for r in data):10.0f}")
print("--- top instructions by samples")
for r in sorted(data, key=lambda r: -f(r, '# Samples'))[:topn]:
    top = sorted(stalls, key=lambda s: -f(r, s))[:2]
    print(r[ix['Address']][-5:], f"{f(r, '# Samples'):6.0f}", f"{f(r, 'Instructions Executed'):11.0f}", r[ix['Source']][:64],
          [(t[6:], int(f(r, t))) for t in top])
