#!/usr/bin/env python
"""Opcode mix and the most-sampled SASS instructions of one profiled launch.

    ncu -i X.ncu-rep --page source --csv > src.csv ; ncu -i X.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_sass_mix.py src.csv raw.csv [n_top]
"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    h = rows[1]
    isrc, ie, iss = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
    stall = {k: h.index(k) for k in ('stall_long_sb', 'stall_barrier', 'stall_short_sb', 'stall_wait',
                                     'stall_not_selected', 'stall_math', 'stall_mio', 'stall_lg')}
    ops, samp, tot, tots, data = collections.Counter(), collections.Counter(), 0, 0, []
    for r in rows[2:]:
        try:
            e, s = int(r[ie]), int(r[iss])
        except Exception:
            continue
        op = [o for o in r[isrc].split() if not o.startswith('@')]
        name = op[0].split('.')[0] if op else '?'
        ops[name] += e; samp[name] += s; tot += e; tots += s
        data.append((e, s, r[isrc].strip(), {k: int(r[i]) for k, i in stall.items()}))
    print(f"warp instructions executed {tot}, stall samples {tots}")
    for k, v in ops.most_common(24):
        print(f"  {k:10s} {v:12d} {100 * v / tot:5.1f} %   samples {100 * samp[k] / tots:5.1f} %")
    n_top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    print("most-sampled instructions:")
    for e, s, src, st in sorted(data, key=lambda x: -x[1])[:n_top]:
        why = max(st, key=st.get)
        print(f"  {100 * s / tots:5.1f} %  exec {e:10d}  {why:18s} | {src[:80]}")
    if len(sys.argv) > 2:
        rr = list(csv.reader(open(sys.argv[2])))
        d = dict(zip(rr[0], rr[2]))
        keys = [k for k in rr[0] if 'issue_stalled' in k and k.endswith('per_issue_active.ratio')]
        print("warp stall reasons (warps per issue-active cycle):")
        for v, k in sorted(((float(d[k] or 0), k) for k in keys), reverse=True)[:8]:
            print(f"  {v:7.3f} {k.split('issue_stalled_')[1].split('_per_issue')[0]}")
        for k in ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.avg.per_cycle_elapsed',
                  'smsp__warps_eligible.avg.per_cycle_active', 'gpu__time_duration.sum',
                  'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct'):
            print(f"  {k} = {d.get(k)}")


if __name__ == '__main__':
    main()
