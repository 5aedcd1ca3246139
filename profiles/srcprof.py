"""Summarise `ncu --page source --print-source cuda,sass --csv` output by CUDA source line (samples, instructions)."""
import csv
import sys
from collections import defaultdict


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    fname = None
    hdr = None
    agg = defaultdict(lambda: [0, 0, ''])
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            fname = r[1].split('/')[-1]
            continue
        if r[0] == 'Line No':
            hdr = r
            si = hdr.index('# Samples'); ie = hdr.index('Instructions Executed')
            continue
        if hdr is None or r[0] in ('Function Name',):
            continue
        if r[0] not in ('', None) and r[0].isdigit():   # a CUDA source line row (aggregated)
            try:
                s = int(r[si]); n = int(r[ie])
            except ValueError:
                continue
            k = (fname, int(r[0]))
            agg[k][0] += s; agg[k][1] += n; agg[k][2] = r[1].strip()[:100]
    tot = sum(v[0] for v in agg.values()) or 1
    toti = sum(v[1] for v in agg.values()) or 1
    print('total samples %d, warp instructions %d' % (tot, toti))
    for (f, l), (s, n, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print('%5.1f%% smp %5.1f%% ins  %s:%d  %s' % (100.0 * s / tot, 100.0 * n / toti, f, l, src))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
