"""Summarise an .ncu-rep (ncu --set full) into a small CSV + JSON under profiles/:
    python tools/ncu_summary.py gpurun_out/gather_v1.ncu-rep profiles/r01_gather_v1 [config-name]
"""
import csv, io, json, os, subprocess, sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__warps_eligible.avg.per_cycle_active']


def to_bytes(v, unit):
    v = float(v.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    cfg = sys.argv[3] if len(sys.argv) > 3 else None
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    summary = []
    for r in rows[2:]:
        d = {'kernel': r[idx['Kernel Name']].split('(')[0].replace('void ', '').replace('<unnamed>::', '').strip()}
        for k in KEYS:
            if k in idx:
                d[k] = r[idx[k]] + ' ' + units[idx[k]]
        d['dram_bytes_total'] = to_bytes(r[idx['dram__bytes_read.sum']], units[idx['dram__bytes_read.sum']]) + \
            to_bytes(r[idx['dram__bytes_write.sum']], units[idx['dram__bytes_write.sum']])
        summary.append(d)
    os.makedirs(os.path.dirname(out) or '.', exist_ok=True)
    with open(out + '.json', 'w') as f:
        json.dump(summary, f, indent=1)
    with open(out + '.txt', 'w') as f:
        for d in summary:
            f.write('== %s\n' % d['kernel'])
            for k, v in d.items():
                if k != 'kernel':
                    f.write('   %-80s %s\n' % (k, v))
    if cfg:
        tp = os.path.join(os.path.dirname(out) or '.', 'traffic_latest.json')
        t = json.load(open(tp)) if os.path.exists(tp) else {}
        e = t.setdefault(cfg, {})
        for d in summary:
            kn = d['kernel']
            if 'density' in kn or kn.endswith(', 0>'): name = 'density'
            elif kn.endswith(', 1>') or 'pressure' in kn: name = 'pressure'
            elif kn.endswith(', 2>') or 'viscosity' in kn: name = 'viscosity'
            else: name = kn.split('<')[0].split('::')[-1].replace('k_', '')
            e[name] = d['dram_bytes_total']
        json.dump(t, open(tp, 'w'), indent=1)
    print(open(out + '.txt').read())


if __name__ == '__main__':
    main()
