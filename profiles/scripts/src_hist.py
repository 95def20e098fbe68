"""Per-source-line instruction histogram of one kernel from an `ncu --page source --csv` export and the current build's line table.
usage: python profiles/scripts/src_hist.py <source.csv> <cubin> <mangled-name-fragment> [min_share]"""
import csv, re, subprocess, os, sys
csvf, cubin, frag = sys.argv[1:4]
thr = float(sys.argv[4]) if len(sys.argv) > 4 else 0.004
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
hdr = rows[hi]; ie = hdr.index('Instructions Executed'); ad = hdr.index('Address'); src = hdr.index('Source')
per = {}; ops = {}; base = None; tot = 0
for r in rows[hi + 1:]:
    if len(r) <= ie or not r[ad]: continue
    try: a = int(r[ad], 16) if r[ad].startswith('0x') else int(r[ad])
    except ValueError: break
    if base is None: base = a
    v = float(r[ie] or 0)
    line = amap.get(a - base, ('?', -1)); per[line] = per.get(line, 0) + v; tot += v
    t = r[src].split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    ops.setdefault(line, {}); ops[line][op] = ops[line].get(op, 0) + v
print('%.1f M warp instructions' % (tot / 1e6))
srcs = {}
for (f, l), v in sorted(per.items()):
    if v / tot < thr: continue
    try:
        if f not in srcs: srcs[f] = open('/root/repo/orb-ygz-slam_b200/csrc/' + f).read().split('\n')
        s = srcs[f][l - 1].strip()[:80]
    except Exception: s = '?'
    top = sorted(ops[(f, l)].items(), key=lambda x: -x[1])[:5]
    print('%5.1f%% %s:%d %s   [%s]' % (100 * v / tot, f[:16], l, s, ' '.join('%s %.1f' % (k, 100 * x / tot) for k, x in top)))
