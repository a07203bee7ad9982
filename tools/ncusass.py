#!/usr/bin/env python3
"""Aggregate the SASS source page of an .ncu-rep: executed warp instructions and stall samples per opcode."""
import csv
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    per_op = defaultdict(lambda: [0, 0, 0.0])
    instrs = []
    for r in rows:
        if r and r[0] == 'Address':
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        src = d['Source'].strip()
        toks = src.split()
        op = toks[1] if toks and toks[0].startswith('@') and len(toks) > 1 else (toks[0] if toks else '?')
        op = op.split('.')[0]
        ex = int(d['Instructions Executed'] or 0)
        smp = int(d['# Samples'] or 0)
        thr = float(d['Avg. Threads Executed'] or 0)
        per_op[op][0] += ex
        per_op[op][1] += smp
        per_op[op][2] += ex * thr
        stalls = {k[6:]: int(v or 0) for k, v in d.items() if k.startswith('stall_') and 'Not Issued' not in k}
        instrs.append((smp, ex, src, stalls, d.get('L1 Wavefronts Shared', '0'), d.get('L1 Wavefronts Shared Ideal', '0')))
    tot_ex = sum(v[0] for v in per_op.values())
    tot_s = sum(v[1] for v in per_op.values())
    print('total warp instructions %d, samples %d' % (tot_ex, tot_s))
    print('%-10s %14s %6s %10s %6s %6s' % ('opcode', 'executed', '%', 'samples', '%', 'thr'))
    for op, (ex, smp, thr) in sorted(per_op.items(), key=lambda kv: -kv[1][0])[:topn]:
        print('%-10s %14d %6.1f %10d %6.1f %6.1f' % (op, ex, 100.0 * ex / max(tot_ex, 1), smp, 100.0 * smp / max(tot_s, 1), thr / max(ex, 1)))
    print('--- hottest instructions by samples')
    for smp, ex, src, stalls, w, wi in sorted(instrs, key=lambda t: -t[0])[:topn]:
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:2]
        print('%7d %10d  %-60s %s  smem_wf %s/%s' % (smp, ex, src[:60], ' '.join('%s=%d' % kv for kv in top if kv[1]), w, wi))
    agg = defaultdict(int)
    for smp, ex, src, stalls, w, wi in instrs:
        for k, v in stalls.items():
            agg[k] += v
    print('--- stall reasons:', ' '.join('%s=%.1f%%' % (k, 100.0 * v / max(tot_s, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))


if __name__ == '__main__':
    main()
