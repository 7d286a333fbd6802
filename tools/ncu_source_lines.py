"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA-C source line.
usage: python tools/ncu_source_lines.py export.csv [top_n]"""
import collections
import csv
import sys


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    hdr = None
    cur_file = cur_line = None
    src = {}
    agg, samp, shw = collections.Counter(), collections.Counter(), collections.Counter()
    seen = set()
    tot = 0
    for r in rows:
        if len(r) == 2 and r[0] == 'File Path':
            cur_file = r[1].split('/')[-1]
            continue
        if r and r[0] == 'Line No':
            hdr = r
            i_inst, i_smp = hdr.index('Instructions Executed'), hdr.index('# Samples')
            i_wf = hdr.index('L1 Wavefronts Shared') if 'L1 Wavefronts Shared' in hdr else None
            continue
        if hdr is None or len(r) < 10:
            continue
        if r[0].isdigit():
            cur_line = (cur_file, int(r[0]))
            src[cur_line] = r[1].strip()[:100]
            continue
        if r[2] in seen:       # every SASS row is listed twice
            continue
        seen.add(r[2])
        n = num(r[i_inst])
        agg[cur_line] += n
        samp[cur_line] += num(r[i_smp])
        shw[cur_line] += num(r[i_wf]) if i_wf is not None else 0
        tot += n
    print('total warp instructions', tot, ' stall samples', sum(samp.values()), ' shared wavefronts', sum(shw.values()))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
        print('%-20s %10d %5.1f%%  smp %5d  wf %9d | %s' % ('%s:%d' % k, v, 100.0 * v / tot, samp[k], shw[k], src.get(k, '')))


if __name__ == '__main__':
    main()
