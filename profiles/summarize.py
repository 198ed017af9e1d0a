#!/usr/bin/env python
"""Turn the ncu artefacts that gpurun brought back (gpurun_out/) into the small tracked summaries
under profiles/:  `python profiles/summarize.py <tag> [--launches csv] [--rep kernel=file.ncu-rep ...]`.

  profiles/<tag>_launches.csv      the per-launch list (gpu__time_duration.sum, --clock-control none)
  profiles/<tag>_launches.md       per-kernel totals and each kernel's SHARE of the profiled command
  profiles/<tag>_<kernel>.json/.md key metrics of one `ncu --set full` capture (read with `ncu -i`)
  profiles/roofline_traffic.json   dram bytes per launch of the dominant kernel (bench.py reads it)
"""
import argparse
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

KEYS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
    'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
    'sass__inst_executed_shared_loads', 'sass__inst_executed_shared_stores', 'sass__inst_executed_global_loads',
    'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum',
    'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed',
    'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed', 'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed',
    'smsp__thread_inst_executed_per_inst_executed.ratio',
]


def fp64_thread_inst(out):
    """Executed fp64 arithmetic thread-instructions of the launch: (DFMA + DMUL + DADD per elapsed cycle, summed over
    the SM sub-partitions) x elapsed cycles -- the `--set full` capture reports the SASS op counters per cycle."""
    try:
        per_cycle = sum(out['smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed' % op]['value'] for op in ('dfma', 'dmul', 'dadd'))
        return per_cycle * out['sm__cycles_elapsed.avg']['value']
    except KeyError:
        return None


def launches(tag, path):
    dst = os.path.join(HERE, tag + '_launches.csv')
    if os.path.abspath(path) != dst:
        shutil.copyfile(path, dst)
    rows = list(csv.reader(open(dst)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    h = rows[hi]
    kn, mv, mu = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(',', ''))
        v = {'ns': v * 1e-3, 'us': v, 'ms': v * 1e3, 'second': v * 1e6}.get(r[mu], v)
        agg.setdefault(r[kn], []).append(v)
    total = sum(sum(v) for v in agg.values())
    with open(os.path.join(HERE, tag + '_launches.md'), 'w') as f:
        f.write('# %s: per-kernel device time of one profiled bench.py command\n\n' % tag)
        f.write('Source: `%s_launches.csv` (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised '
                'launches: compare SHARES).\n\n| kernel | launches | total us | mean us | share |\n|---|---:|---:|---:|---:|\n' % tag)
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write('| `%s` | %d | %.1f | %.1f | %.1f%% |\n' % (k[:90], len(v), sum(v), sum(v) / len(v), 100 * sum(v) / total))
    print('wrote', tag + '_launches.md')


def capture(tag, name, rep, traffic_key=None, points=0):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units, vals = rows[0], rows[1], rows[2]
    col = {n: (u, v) for n, u, v in zip(h, units, vals)}
    out = collections.OrderedDict()
    out['kernel'] = col.get('Kernel Name', ('', ''))[1]
    for k in KEYS:
        if k in col:
            u, v = col[k]
            try:
                out[k] = {'value': float(v.replace(',', '')), 'unit': u}
            except ValueError:
                out[k] = {'value': v, 'unit': u}
    json.dump(out, open(os.path.join(HERE, '%s_%s.json' % (tag, name)), 'w'), indent=1)
    with open(os.path.join(HERE, '%s_%s.md' % (tag, name)), 'w') as f:
        f.write('# %s: ncu --set full capture of `%s`\n\n| metric | value | unit |\n|---|---:|---|\n' % (tag, out['kernel'][:100]))
        for k, v in out.items():
            if k != 'kernel':
                f.write('| %s | %s | %s |\n' % (k, v['value'], v['unit']))
    if traffic_key:
        def mb(k):
            v = out[k]
            return v['value'] * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[v['unit']]
        tpath = os.path.join(HERE, 'roofline_traffic.json')
        t = json.load(open(tpath)) if os.path.exists(tpath) else {}
        t[traffic_key] = mb('dram__bytes_read.sum') + mb('dram__bytes_write.sum')
        t[traffic_key + '_source'] = '%s_%s.json' % (tag, name)
        fp = out.get('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed')
        if fp:
            t[traffic_key.replace('bytes_per_launch', 'fp64_pipe_pct')] = fp['value']
        if points and 'smsp__inst_executed.sum' in out:
            t[traffic_key.replace('bytes_per_launch', 'warp_inst_per_point')] = out['smsp__inst_executed.sum']['value'] / (points / 32.0)
        inst = fp64_thread_inst(out)
        if inst and points:
            t[traffic_key.replace('bytes_per_launch', 'fp64_thread_inst_per_point')] = inst / points
            t[traffic_key.replace('bytes_per_launch', 'fp64_inst_source')] = '%s_%s.json: (dfma + dmul + dadd thread instructions per cycle) x cycles / %d points' % (tag, name, points)
            for op in ('dfma', 'dmul', 'dadd'):
                t[traffic_key.replace('bytes_per_launch', op + '_per_point')] = \
                    out['smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed' % op]['value'] * out['sm__cycles_elapsed.avg']['value'] / points
        json.dump(t, open(tpath, 'w'), indent=1)
    print('wrote', '%s_%s.md' % (tag, name))


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('tag')
    ap.add_argument('--launches')
    ap.add_argument('--rep', action='append', default=[], help='name=path[:traffic_key[:points per launch]]')
    a = ap.parse_args()
    if a.launches:
        launches(a.tag, a.launches)
    for spec in a.rep:
        name, rest = spec.split('=', 1)
        path, _, key = rest.partition(':')
        key, _, pts = key.partition(':')
        capture(a.tag, name, path, key or None, int(pts or 0))
