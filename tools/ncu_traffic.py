#!/usr/bin/env python
"""profiles/traffic.json entry from one `ncu --set full` capture: DRAM bytes read + written per launch.

  python tools/ncu_traffic.py gpurun_out/prof_leduc-holdem_r01g.ncu-rep leduc-holdem:uint8:65536:128
"""
import csv
import json
import os
import subprocess
import sys

UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}


def main():
    rep, key = sys.argv[1], sys.argv[2]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    tot = 0.0
    for name in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        i = hdr.index(name)
        tot += sum(float(r[i].replace(',', '')) * UNIT[units[i]] for r in data) / len(data)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'traffic.json')
    d = json.load(open(path)) if os.path.exists(path) else {}
    d[key] = tot
    json.dump(d, open(path, 'w'), indent=1, sort_keys=True)
    print(key, '%.1f MB per launch' % (tot / 1e6))


if __name__ == '__main__':
    main()
