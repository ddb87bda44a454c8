"""Turns the raw ncu captures brought back in gpurun_out/ into the small text summaries committed under profiles/.
Usage: python profiles/summarize.py <round-tag>   (e.g. r1b)"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__inst_executed_pipe_fp64.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size', 'launch__block_size',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
STALLS = 'smsp__average_warps_issue_stalled_'


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main(tag):
    for model in ('three_circle', 'circular'):
        rep = 'gpurun_out/prof_%s_%s.ncu-rep' % (model, tag)
        hdr, units, rows = raw(rep)
        lines = ['# ncu --set full --clock-control none, the kernels of the step (%s), 1M agents, 1 agent/m^2; one section per captured launch'
                 % model, '# source: %s (not committed: binary); metric, unit, value' % rep]
        for r in rows:
            d = dict(zip(hdr, r))
            lines.append('kernel: %s' % d.get('Kernel Name'))
            for k in KEYS:
                if k in d:
                    lines.append('%-75s %-16s %s' % (k, units[hdr.index(k)], d[k]))
            st = sorted(((float(d[h]), h) for h in hdr if h.startswith(STALLS) and h.endswith('per_issue_active.ratio') and d[h]),
                        reverse=True)
            lines.append('stall reasons (warps per issue-active cycle):')
            for v, h in st[:8]:
                lines.append('    %-40s %.3f' % (h[len(STALLS):].replace('_per_issue_active.ratio', ''), v))
        open('profiles/ncu_full_%s_%s.txt' % (model, tag), 'w').write('\n'.join(lines) + '\n')
        # launch list: per-kernel totals over the profiled launches
        agg = {}
        import os
        csv_path = 'gpurun_out/launches_%s_%s.csv' % (model, tag)
        if not os.path.exists(csv_path):
            csv_path = 'gpurun_out/launches_%s_%s.csv' % (model, tag[:2])
        with open(csv_path) as f:
            rd = csv.reader(l for l in f if not l.startswith('=='))
            h = next(rd)
            ki, vi, mi = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Name')
            for r in rd:
                if len(r) <= vi or r[mi] != 'gpu__time_duration.sum':
                    continue
                name = r[ki].split('(')[0]
                a = agg.setdefault(name, [0, 0.0])
                a[0] += 1
                a[1] += float(r[vi].replace(',', ''))
        tot = sum(a[1] for a in agg.values())
        lines = ['# ncu --metrics gpu__time_duration.sum --clock-control none: all launches of one short `bench.py --model %s` run (see capture_*.sh)' % model,
                 '# (cold-cache, serialised: compare SHARES; includes upload/download kernels of the e2e leg)',
                 '%-40s %8s %14s %8s' % ('kernel', 'launches', 'total ns', 'share')]
        for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            lines.append('%-40s %8d %14.0f %7.1f%%' % (name, c, t, 100 * t / tot))
        open('profiles/launches_%s_%s.txt' % (model, tag), 'w').write('\n'.join(lines) + '\n')


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'r1b')
