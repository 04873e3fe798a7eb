"""Summarise an ncu `--page source --csv --print-source cuda,sass` dump per CUDA source line.

usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv; python tools/ncu_lines.py src.csv [kernel-substring]
"""
import csv
import sys
from collections import defaultdict


def main(path, kernel=None, top=45):
    rows = list(csv.reader(open(path)))
    cur_file, cur_fn = None, None
    agg = defaultdict(lambda: [0, 0, ""])   # (fn, file, line) -> [inst, samples, text]
    hdr = None
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            cur_file = r[1].split('/')[-1]
            continue
        if r[0] == 'Function Name':
            cur_fn = r[1]
            continue
        if r[0] == 'Line No':
            hdr = r
            continue
        if hdr is None or r[0] == '' or not r[0].isdigit():
            continue
        try:
            inst = int(r[hdr.index('Instructions Executed')])
            samp = int(r[hdr.index('# Samples')])
        except (ValueError, IndexError):
            continue
        if kernel and kernel not in cur_fn:
            continue
        key = (cur_fn[:60], cur_file, int(r[0]))
        agg[key][0] += inst
        agg[key][1] += samp
        agg[key][2] = r[1][:90]
    tot_i = sum(v[0] for v in agg.values())
    tot_s = sum(v[1] for v in agg.values())
    print("total warp-instructions %d, samples %d" % (tot_i, tot_s))
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%5.1f%% inst %5.1f%% smp  %-18s:%4d  %s" % (100.0 * v[0] / max(tot_i, 1), 100.0 * v[1] / max(tot_s, 1),
                                                          key[1], key[2], v[2]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
