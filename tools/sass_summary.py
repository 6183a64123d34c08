"""Static SASS instruction counts per kernel family of libraleigh_b200.so (cuobjdump -sass), as a Markdown table.
    python tools/sass_summary.py > profiles/sass_summary.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'raleigh_b200', 'libraleigh_b200.so')
out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
COLS = [('UTCHMMA', r'\bUTCHMMA'), ('LDTM', r'\bLDTM'), ('UTMALDG', r'\bUTMALDG'), ('UBLKCP', r'\bUBLKCP'),
        ('SYNCS', r'\bSYNCS'), ('DMMA', r'\bDMMA'), ('DFMA', r'\bDFMA'), ('FFMA', r'\bFFMA'), ('UCGABAR', r'\bUCGABAR'),
        ('LDGSTS', r'\bLDGSTS'), ('LDG.128', r'\bLDG\.E\.(?:\w+\.)*128'), ('LDS.128', r'\bLDS\.128'), ('SHFL', r'\bSHFL'),
        ('RSQ64H', r'RSQ64H'), ('BAR', r'\bBAR\.')]
fam = collections.defaultdict(lambda: {'inst': 0, **{c: 0 for c, _ in COLS}})
total = {c: 0 for c, _ in COLS}
nk = 0
cur = None
counts = None


def flush():
    global cur, counts
    if cur is None:
        return
    f = fam[cur]
    f['inst'] += 1
    for c, _ in COLS:
        f[c] = max(f[c], counts[c])
        total[c] += counts[c]


for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        flush()
        nk += 1
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r'<.*', '', name.replace('void ', ''))
        name = re.sub(r'\(.*', '', name)
        cur, counts = name, {c: 0 for c, _ in COLS}
        continue
    if cur is None:
        continue
    for c, pat in COLS:
        if re.search(pat, line):
            counts[c] += 1
flush()
print('# SASS summary of `libraleigh_b200.so` (round 2, final tree; `python tools/sass_summary.py`: `cuobjdump -sass`, sm_100a)\n')
print('Static instruction counts per kernel family (maximum over the template instances of the family; `inst` = number of instances in the library).')
print('`UTCHMMA` = `tcgen05.mma`, `LDTM` = `tcgen05.ld` (TMEM read-back), `UTMALDG` = `cp.async.bulk.tensor` (TMA), `UBLKCP` = `cp.async.bulk` (non-tensor bulk copy), '
      '`SYNCS` = mbarrier operations, `DMMA` = `mma.sync.m8n8k4.f64`, `LDGSTS` = `cp.async`, `UCGABAR` = `barrier.cluster`, `RSQ64H` = FP64 rsqrt seed.\n')
print('| kernel family | inst | ' + ' | '.join(c for c, _ in COLS) + ' |')
print('|---|---:|' + '---:|' * len(COLS))
keep = [k for k, v in fam.items() if any(v[c] for c in ('UTCHMMA', 'UTMALDG', 'UBLKCP', 'DMMA', 'UCGABAR', 'LDGSTS')) or v['DFMA'] + v['FFMA'] >= 60]
for k in sorted(keep, key=lambda k: (-fam[k]['UTCHMMA'], -fam[k]['UTMALDG'], -fam[k]['UBLKCP'], -fam[k]['DMMA'], k)):
    v = fam[k]
    print('| `%s` | %d | ' % (k, v['inst']) + ' | '.join(str(v[c]) if v[c] else '·' for c, _ in COLS) + ' |')
print('\nWhole library (%d kernels): ' % nk + ', '.join('%s %d' % (c, total[c]) for c, _ in COLS) + '.')
