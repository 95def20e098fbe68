"""Per-source-line stall-sample histogram (and instruction counts) of one kernel from an `ncu --page source --csv` export.
usage: python profiles/scripts/stall_hist.py <source.csv> <cubin> <mangled-name-fragment> [min_share]"""
import csv, re, subprocess, os, sys
csvf, cubin, frag = sys.argv[1:4]
thr = float(sys.argv[4]) if len(sys.argv) > 4 else 0.01
out = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
txt = out.split('\n')
start = [i for i, l in enumerate(txt) if '.text.' in l and frag in l and '.section' in l][0]
cur = None; amap = {}
for l in txt[start + 1:]:
    if l.startswith('//---------------------') and amap: break
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
    if m and cur is not None: amap[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(csvf)))
hi = [i for i, r in enumerate(rows) if 'Instructions Executed' in r][0]
hdr = rows[hi]; ie = hdr.index('Instructions Executed'); ad = hdr.index('Address'); st = hdr.index('Warp Stall Sampling (All Samples)')
reasons = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
per = {}; pins = {}; rs = {}; base = None; tot = 0; toti = 0; rtot = {}
for r in rows[hi + 1:]:
    if len(r) <= max(reasons) or not r[ad]: continue
    try: a = int(r[ad], 16) if r[ad].startswith('0x') else int(r[ad])
    except ValueError: break
    if base is None: base = a
    v = float(r[st] or 0); vi = float(r[ie] or 0)
    line = amap.get(a - base, ('?', -1))
    per[line] = per.get(line, 0) + v; pins[line] = pins.get(line, 0) + vi; tot += v; toti += vi
    d = rs.setdefault(line, {})
    for i in reasons:
        x = float(r[i] or 0)
        if x: d[hdr[i]] = d.get(hdr[i], 0) + x; rtot[hdr[i]] = rtot.get(hdr[i], 0) + x
print('%.2f M warp instructions, %d stall samples; reasons: %s' % (toti / 1e6, tot, ' '.join('%s %.0f%%' % (k[6:], 100 * v / sum(rtot.values())) for k, v in sorted(rtot.items(), key=lambda x: -x[1])[:8])))
srcs = {}
for (f, l), v in sorted(per.items()):
    if v / tot < thr: continue
    try:
        if f not in srcs: srcs[f] = open('/root/repo/orb-ygz-slam_b200/csrc/' + f).read().split('\n')
        s = srcs[f][l - 1].strip()[:70]
    except Exception: s = '?'
    top = sorted(rs[(f, l)].items(), key=lambda x: -x[1])[:3]
    print('%5.1f%% stalls %5.1f%% instr  %s:%d %s   [%s]' % (100 * v / tot, 100 * pins[(f, l)] / toti, f[:14], l, s, ' '.join('%s %.1f' % (k[6:], 100 * x / tot) for k, x in top)))
