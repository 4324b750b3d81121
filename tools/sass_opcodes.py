"""Opcode histogram per kernel of the built library (cuobjdump -sass): the evidence that the hot kernels use tcgen05 / TMEM,
bulk-async copies, mbarriers, thread-block clusters and programmatic dependent launch.  No GPU needed.
Usage: python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'bruno_b200', 'lib', 'libbruno_sm100.so')
sass = subprocess.run('cuobjdump -sass %s | c++filt' % so, shell=True, capture_output=True, text=True).stdout
COLS = ['UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UBLKCP.S.G', 'UBLKCP.G.S', 'STAS', 'UCGABAR_ARV', 'PREEXIT', 'ACQBULK', 'SYNCS', 'MUFU']
kernels, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r'\s*Function : (.*)', line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_.]+)', line)
    if m and cur:
        op = m.group(1)
        kernels[cur]['total'] += 1
        for c in COLS:
            if op == c or op.startswith(c + '.') or (c in ('UBLKCP.S.G', 'UBLKCP.G.S') and op.startswith(c)):
                kernels[cur][c] += 1
print('# cuobjdump -sass bruno_b200/lib/libbruno_sm100.so, opcode counts per kernel (tools/sass_opcodes.py).  tcgen05: UTCHMMA = tcgen05.mma,')
print('# UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st; TMA engine: UBLKCP = cp.async.bulk (1-D bulk copies on pre-swizzled slabs: no')
print('# tensor maps, hence no UTMALDG / UTMASTG); STAS = st.async.shared::cluster (split-K partial rows pushed to the owner CTA),')
print('# UCGABAR_ARV = barrier.cluster.arrive; PREEXIT / ACQBULK = griddepcontrol.launch_dependents / wait; SYNCS = mbarrier.')
for name, c in sorted(kernels.items(), key=lambda kv: -kv[1]['total']):
    if c['total'] < 200 and not any(c[k] for k in COLS[:10]):
        continue
    short = re.sub(r'\(.*', '', name.replace('(anonymous namespace)::', ''))
    print('%-78s total %6d  ' % (short[:78], c['total']) + ' '.join('%s %d' % (k, c[k]) for k in COLS))
