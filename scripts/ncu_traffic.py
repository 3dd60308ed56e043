#!/usr/bin/env python
"""profiles/ncu_traffic.json from the committed `ncu --set full` raw pages: DRAM bytes of ONE launch
of the dominant kernel per workload (bench.py reports it as roofline.traffic).

    python scripts/ncu_traffic.py profiles/r02_group_pixels_nyuv2_ncu_raw.csv:nyuv2_b8 [...]
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def dram_bytes(path):
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if 'Kernel Name' in r)
    units = rows[rows.index(hdr) + 1]
    data = rows[rows.index(hdr) + 2]
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    total = 0.0
    for name in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        i = hdr.index(name)
        total += float(data[i].replace(',', '')) * scale[units[i]]
    return total, data[hdr.index('Kernel Name')], float(data[hdr.index('gpu__time_duration.sum')].replace(',', ''))


def main():
    out_path = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    table = json.load(open(out_path)) if os.path.exists(out_path) else {}
    for arg in sys.argv[1:]:
        path, key = arg.rsplit(':', 1)
        total, kernel, dur = dram_bytes(path)
        table[key] = {'dram_bytes': total, 'kernel': kernel, 'gpu_time_duration': dur,
                      'source': os.path.relpath(path, ROOT)}
        print(key, f'{total / 1e6:.1f} MB', kernel[:60])
    json.dump(table, open(out_path, 'w'), indent=1, sort_keys=True)


if __name__ == '__main__':
    main()
