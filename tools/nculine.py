#!/usr/bin/env python3
"""Attribute an .ncu-rep's per-SASS-instruction counters to CUDA source lines.

    tools/nculine.py <report.ncu-rep> <library.so> <kernel-name-substring> [top N] [ncu -k filter]

The SASS page of the report lists instructions in address order; nvdisasm -g on the cubin extracted from the
.so lists the same instructions with //## File/line markers (needs -lineinfo).  The .so must be the build that was profiled.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def sass_lines(so, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
    cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith('.cubin')]
    out = []
    for c in cubins:
        txt = subprocess.run(['nvdisasm', '-g', '-c', c], capture_output=True, text=True).stdout
        cur, inside = None, False
        for line in txt.splitlines():
            if line.startswith('//--------------------- .text.'):
                inside = kernel in line
                continue
            if line.startswith('//--------------------- '):
                inside = False
            if not inside:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', line)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
            if m:
                out.append((int(m.group(1), 16), cur, m.group(2).strip()))
    return out


def main():
    rep, so, kernel = sys.argv[1:4]
    topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    lines = sass_lines(so, kernel)
    # a report with several kernels: name the one to read (ncu's -k matches the demangled base name), e.g. regex:k2_filterbank
    kfilter = ['-k', sys.argv[5]] if len(sys.argv) > 5 else []
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'] + kfilter, capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, data = None, []
    for r in rows:
        if r and r[0] == 'Address':
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    if len(data) != len(lines):
        print('warning: %d profiled instructions vs %d in the cubin (different build?)' % (len(data), len(lines)))
    per = defaultdict(lambda: [0, 0, defaultdict(int)])
    tot_e = tot_s = 0
    for d, (addr, loc, text) in zip(data, lines):
        e, s = int(d['Instructions Executed'] or 0), int(d['# Samples'] or 0)
        per[loc][0] += e
        per[loc][1] += s
        for k, v in d.items():
            if k.startswith('stall_') and 'Not Issued' not in k and v and int(v):
                per[loc][2][k[6:]] += int(v)
        tot_e += e
        tot_s += s
    print('total warp instructions %d, samples %d' % (tot_e, tot_s))
    srcs = {}
    for loc, (e, s, st) in sorted(per.items(), key=lambda kv: -kv[1][1])[:topn]:
        text = ''
        if loc:
            path = None
            for root, _, files in os.walk(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'jaadec_b200')):
                if loc[0] in files:
                    path = os.path.join(root, loc[0])
            if path:
                srcs.setdefault(path, open(path).read().splitlines())
                if loc[1] - 1 < len(srcs[path]):
                    text = srcs[path][loc[1] - 1].strip()[:70]
        top = ' '.join('%s=%d%%' % (k, 100 * v // max(s, 1)) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:2])
        print('%5.1f%% smp %5.1f%% ins  %s:%-4s %-70s %s' % (100.0 * s / max(tot_s, 1), 100.0 * e / max(tot_e, 1),
                                                           loc[0][:14] if loc else '?', loc[1] if loc else '', text, top))


if __name__ == '__main__':
    main()
