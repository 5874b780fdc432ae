"""Summarise an .ncu-rep (read with `ncu -i`, no GPU needed): headline metrics + per-code-region stall samples.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [frames] [chunk]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
frames = float(sys.argv[2]) if len(sys.argv) > 2 else None
chunk_n = int(sys.argv[3]) if len(sys.argv) > 3 else 200
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor', 'sm__pipe_tensor_cycles_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio']
for r in rows[2:]:
    print('=' * 100)
    for h, u, v in zip(hdr, units, r):
        if any(h == k or (k.endswith('tensor') or k.endswith('active')) and h.startswith(k) and 'pct' in h and 'avg' in h for k in KEYS):
            print(f'{h} [{u}] = {v}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# several kernels may be concatenated: split on "Kernel Name" rows
blocks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': []}
        blocks.append(cur)
    elif cur is not None:
        cur['rows'].append(r)
for b in blocks:
    hdr, data = b['rows'][0], b['rows'][1:]
    col = {h: i for i, h in enumerate(hdr)}
    g = lambda r, k: int(r[col[k]] or 0) if k in col and r[col[k]] not in ('', None) else 0
    tot = sum(g(r, '# Samples') for r in data) or 1
    totex = sum(g(r, 'Instructions Executed') for r in data) or 1
    print('-' * 100)
    print(b['name'], 'samples', tot, 'inst', totex, ('inst/frame %.0f' % (totex / frames)) if frames else '')
    for start in range(0, len(data), chunk_n):
        ch = data[start:start + chunk_n]
        s = sum(g(r, '# Samples') for r in ch)
        ex = sum(g(r, 'Instructions Executed') for r in ch)
        if ex == 0 and s == 0:
            continue
        ops = {}
        for r in ch:
            t = r[col['Source']].split()
            if not t:
                continue
            op = (t[1] if t[0].startswith('@') and len(t) > 1 else t[0]).split('.')[0]
            ops[op] = ops.get(op, 0) + 1
        top = ' '.join(f'{k}:{v}' for k, v in sorted(ops.items(), key=lambda x: -x[1])[:6])
        st = ' '.join(f'{k[6:]}:{sum(g(r, k) for r in ch)}' for k in ('stall_short_sb', 'stall_wait', 'stall_long_sb', 'stall_branch_resolving',
                                                                     'stall_math', 'stall_mio', 'stall_not_selected', 'stall_no_inst', 'stall_dispatch'))
        per = (' %7.1f/frame' % (ex / frames)) if frames else ''
        print(f'{start:5d}: samples {100 * s / tot:5.1f}%  exec {100 * ex / totex:5.1f}%{per} | {st} | {top}')
