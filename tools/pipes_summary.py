#!/usr/bin/env python3
"""Summarise an ncu multi-metric launch list (tools/gpu_run*.sh) per kernel: device time, share of the step,
fmaheavy (IMAD.WIDE) pipe utilisation, fma / alu instruction mix, DRAM traffic.
  python tools/pipes_summary.py gpurun_out/launches_*.csv [> profiles/....md]"""
import csv
import re
import sys
from collections import defaultdict

SMS, CLK = 148, 1.965e9   # fmaheavy issues 1 warp instruction / 2 cycles / SM sub-partition = 2 per clk per SM


def main(path):
    rows = defaultdict(dict)
    for r in csv.reader(l for l in open(path) if l.startswith('"')):
        if r[0] == 'ID':
            continue
        m = re.match(r'(?:void )?zk_task_kernel<(.*)>\(int', r[4])
        rows[int(r[0])]['name'] = (m.group(1) if m else r[4]).replace('zk::', '')
        rows[int(r[0])][r[12]] = float(r[14].replace(',', ''))
    agg = defaultdict(lambda: defaultdict(float))
    for r in rows.values():
        a = agg[r['name']]
        a['n'] += 1
        for k, v in r.items():
            if k != 'name':
                a[k] += v
    tot = sum(a['gpu__time_duration.sum'] for a in agg.values())
    print(f'| kernel | launches | ms | share | fmaheavy util | fma inst/alu inst | inst/ns | DRAM GB (r+w) |')
    print('|---|---|---|---|---|---|---|---|')
    th = ta = 0.0
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]['gpu__time_duration.sum']):
        ns = a['gpu__time_duration.sum']
        heavy = a['sm__inst_executed_pipe_fmaheavy.sum']
        util = heavy / (ns * 1e-9 * CLK * SMS * 2)
        th += heavy
        print(f"| {name} | {int(a['n'])} | {ns / 1e6:.3f} | {ns / tot:.3f} | {util:.3f} | "
              f"{a['sm__inst_executed_pipe_fma.sum'] / max(a['sm__inst_executed_pipe_alu.sum'], 1):.2f} | "
              f"{a['sm__inst_executed.sum'] / ns:.1f} | {(a['dram__bytes_read.sum'] + a['dram__bytes_write.sum']) / 1e9:.3f} |")
    print(f'\ntotal {tot / 1e6:.2f} ms; whole-list fmaheavy pipe utilisation {th / (tot * 1e-9 * CLK * SMS * 2):.3f} '
          f'(warp instructions on the fmaheavy pipe / (2 per clk per SM x {SMS} SMs x {CLK / 1e9} GHz x time))')


if __name__ == '__main__':
    main(sys.argv[1])
