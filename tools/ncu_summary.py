"""dev helper: summarise an .ncu-rep of the fused kernel (raw metrics + per-role stall samples)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H = rows[0]
want = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_active','lts__throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'l1tex__t_sector_hit_rate.pct','lts__t_sectors_srcunit_tex_op_read.sum','launch__grid_size','launch__block_size','sm__cycles_elapsed.avg',
 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.sum','smsp__inst_executed_pipe_xu.sum']
for w in want:
    if w in H:
        i = H.index(w); print(w, [r[i] for r in rows[1:]])
for i, h in enumerate(H):
    if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio'):
        v = float(rows[2][i])
        if v > 0.1: print('  stall', h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), round(v, 2))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
blk = [r for r in rows[2:] if len(r) > 8]
T = sum(int(r[4]) for r in blk if r[4].isdigit())
cur = ci = 0; start = 0
for i, r in enumerate(blk):
    s = int(r[4]) if r[4].isdigit() else 0
    ie = int(r[5]) if r[5].isdigit() else 0
    cur += s; ci += ie
    if any(m in r[1] for m in ('BAR.SYNC', 'SYNCS.ARRIVE', 'UBLKCP', 'SYNCS.PHASECHK', 'EXIT', 'FENCE.VIEW.ASYNC')):
        if cur > 0.004 * T or 'EXIT' in r[1]:
            print(f"[{start:5d}-{i:5d}] samples={cur:6d} ({100*cur/T:4.1f}%) warp_inst={ci:10d}  {r[1].strip()[:50]}")
            start = i + 1; cur = ci = 0
top = sorted(blk, key=lambda r: -(int(r[4]) if r[4].isdigit() else 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 12]
for r in top: print(r[4], r[5], r[1][:100])
