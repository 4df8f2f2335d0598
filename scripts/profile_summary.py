#!/usr/bin/env python
"""Turn an ncu report of scripts/profile_target.py into the files kept under profiles/:
  <tag>_ncu_full_raw.csv             raw page (every metric of every captured launch)
  <tag>_nuts_kernel_hot_lines.csv    per-source-line stall samples / instructions / shared-memory wavefronts
usage: python scripts/profile_summary.py gpurun_out/prof_X.ncu-rep TAG N_EVALS [LAUNCH]   (LAUNCH: which nuts_kernel launch
of the report the per-line table is for, default 0)"""
import csv, os, subprocess, sys
rep, tag, nev = sys.argv[1], sys.argv[2], float(sys.argv[3])
launch = sys.argv[4] if len(sys.argv) > 4 else '0'
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = os.path.join(root, 'profiles', f'{tag}_ncu_full_raw.csv')
with open(raw, 'w') as fh:
    subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=fh, stderr=subprocess.DEVNULL, check=True)
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:nuts_kernel', '--launch-skip', launch,
                      '--launch-count', '1', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
out, hdr, sec = [], None, 0
for r in csv.reader(src.splitlines()):
    if not r:
        continue
    if r[0] == 'Line No':
        hdr, sec = r, sec + 1
        continue
    if r[0] in ('Kernel Name', 'File Name') or r[0] == '' or hdr is None:
        continue
    try:
        ln = int(r[0]); smp = int(r[6]) if r[6] not in ('-', '') else 0; ins = int(r[7]) if r[7] not in ('-', '') else 0
    except ValueError:
        continue
    out.append((sec, ln, r[1].strip()[:110], smp, ins, r[hdr.index('L1 Wavefronts Shared')], r[hdr.index('L1 Wavefronts Shared Ideal')]))
ts, ti = sum(o[3] for o in out), sum(o[4] for o in out)
out.sort(key=lambda o: -o[3])
with open(os.path.join(root, 'profiles', f'{tag}_nuts_kernel_hot_lines.csv'), 'w') as fh:
    w = csv.writer(fh)
    w.writerow(['file_section', 'line', 'source', 'stall_samples_pct', 'instructions_pct', 'instructions_per_eval',
                'smem_wavefronts', 'smem_wavefronts_ideal'])
    for o in out[:120]:
        w.writerow([o[0], o[1], o[2], round(100 * o[3] / ts, 2), round(100 * o[4] / ti, 2), round(o[4] / nev, 1), o[5], o[6]])
print(f'{tag}: {ti / nev:.0f} warp instructions per evaluation')
for o in out[:12]:
    print(f'  {o[1]:4d} samples {100 * o[3] / ts:5.2f}% instr {100 * o[4] / ti:5.2f}% | {o[2][:90]}')
rows = list(csv.reader(open(raw)))
idx = {h: i for i, h in enumerate(rows[0])}
keys = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__grid_size',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio']
seen = set()
for r in rows[2:]:
    name = r[idx['Kernel Name']].split('(')[0]
    if name in seen:
        continue
    seen.add(name)
    print(name)
    for k in keys:
        if k in idx:
            print(f'    {k:85s} {r[idx[k]]} {rows[1][idx[k]]}')
