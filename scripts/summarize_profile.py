"""Turn the ncu artefacts of a GPU visit (gpurun_out/) into the tracked summaries under profiles/.

    python scripts/summarize_profile.py r01        # reads gpurun_out/launches.csv, gpurun_out/prof_*.ncu-rep

Writes profiles/<round>_launches.csv (kernel, grid, block, duration), profiles/<round>_ncu_<name>.csv (selected raw
metrics of each full-set capture) and refreshes profiles/ncu_summary.json (per-launch DRAM traffic that bench.py
reports as roofline.traffic)."""
import csv
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'profiles')
SRC = os.path.join(ROOT, 'gpurun_out')
KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__cycles_elapsed.max',
        'smsp__cycles_active.avg', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__cluster_max_active', 'launch__grid_size', 'launch__block_size',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_bytes.sum',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio']


def launches(tag):
    """launches.csv: two eager passes (scripts/prof_one.py); launches_bench.csv: the bench.py command itself."""
    for src, dst in (('launches.csv', f'{tag}_launches.csv'), ('launches_bench.csv', f'{tag}_launches_bench.csv')):
        path = os.path.join(SRC, src)
        if not os.path.exists(path):
            continue
        lines = [l for l in open(path) if not l.startswith('==')]
        rows = list(csv.DictReader(lines))
        share = {}
        with open(os.path.join(OUT, dst), 'w', newline='') as f:
            w = csv.writer(f)
            w.writerow(['id', 'kernel', 'grid', 'block', 'duration_ns'])
            for r in rows:
                name = r['Kernel Name'].replace('mp::<unnamed>::', '')
                w.writerow([r['ID'], name, r['Grid Size'], r['Block Size'], r['Metric Value']])
                key = name.split('(')[0].replace('void ', '')
                share[key] = share.get(key, 0.0) + float(r['Metric Value'])
        tot = sum(share.values()) or 1.0
        print('wrote', dst, len(rows), 'launches; time share by kernel:')
        for k, v in sorted(share.items(), key=lambda kv: -kv[1])[:10]:
            print(f'    {100 * v / tot:5.1f}%  {k[:90]}')


def full_sets(tag):
    summary_path = os.path.join(OUT, 'ncu_summary.json')
    summary = json.load(open(summary_path)) if os.path.exists(summary_path) else {}
    for rep in sorted(glob.glob(os.path.join(SRC, 'prof_*.ncu-rep'))):
        name = os.path.basename(rep)[5:-8]
        raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        with open(os.path.join(OUT, f'{tag}_ncu_{name}.csv'), 'w', newline='') as f:
            w = csv.writer(f)
            w.writerow(['kernel', 'metric', 'unit', 'value'])
            for r in rows[2:]:
                kname = r[idx['Kernel Name']]
                for m in KEEP:
                    if m in idx:
                        w.writerow([kname, m, units[idx[m]], r[idx[m]]])
        r = rows[2]
        def val(m):
            v, u = float(r[idx[m]]), units[idx[m]].lower()
            return v * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1)
        key = {'rec_b256': 'lstm_rec_h256', 'rec_b1': 'lstm_rec_h256_b1', 'gemm_tc': 'gemm_tf32x3',
               'rec_tc_b256': 'lstm_rec_tc_h256', 'gemm_ffma2': 'gemm_linear', 'rec_h64_rows': 'lstm_rec_h64',
               'rec_f16_b256': 'lstm_rec_f16_h256_tile64', 'rec_f16w_b256': 'lstm_rec_f16_h256', 'gemm_f16': 'gemm_f16x3',
               'gemm_f16_linear1': 'gemm_f16x3_linear1', 'gemm_f16_k512': 'gemm_f16x3_k512'}.get(name, name)
        summary[key] = {'round': tag, 'kernel': r[idx['Kernel Name']], 'grid': r[idx['Grid Size']],
                        'dram_bytes_per_launch': val('dram__bytes_read.sum') + val('dram__bytes_write.sum'),
                        'duration_ms_under_ncu': float(r[idx['gpu__time_duration.sum']]) * {'us': 1e-3, 'ms': 1, 'ns': 1e-6, 's': 1e3}.get(units[idx['gpu__time_duration.sum']].replace('second', 's'), 1)}
        print('wrote', f'{tag}_ncu_{name}.csv', summary[key])
    json.dump(summary, open(summary_path, 'w'), indent=1, sort_keys=True)


if __name__ == '__main__':
    tag = sys.argv[1] if len(sys.argv) > 1 else 'r01'
    os.makedirs(OUT, exist_ok=True)
    launches(tag)
    full_sets(tag)
