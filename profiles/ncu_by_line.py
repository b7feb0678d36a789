"""Join ncu's per-SASS-instruction counts with nvdisasm line info -> executed warp instructions per
source line / phase.  Usage: python profiles/ncu_by_line.py REP.ncu-rep LIB.so KERNEL_MANGLED_SUBSTR"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, lib, kname = sys.argv[1:4]
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
# locate function (the library holds one cubin per translation unit)
for cub in [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith('.cubin')]:
    dis = subprocess.run(['nvdisasm', '-g', '-c', cub], capture_output=True, text=True).stdout.splitlines()
    start = next((i for i, l in enumerate(dis) if l.startswith('.text.') and kname in l), None)
    if start is not None:
        break
lines = {}   # offset -> (file line)
cur = None
for l in dis[start + 1:]:
    if l.startswith('//-----') and '.text.' in l:
        break
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        lines[int(m.group(1), 16)] = cur
rows = list(csv.reader(io.StringIO(subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
base = None
byline = collections.Counter(); thr = collections.Counter(); samples = collections.Counter(); tot = 0
for r in rows[2:]:
    try:
        addr = int(r[ix['Address']], 16) if r[ix['Address']].startswith('0x') else int(r[ix['Address']])
        n = int(r[ix['Instructions Executed']]); t = int(r[ix['Thread Instructions Executed']]); sm = int(r[ix['# Samples']])
    except (ValueError, IndexError):
        continue
    if base is None:
        base = addr
    key = lines.get(addr - base)
    byline[key] += n; thr[key] += t; samples[key] += sm; tot += n
stot = sum(samples.values())
print(f'total warp instr {tot}, samples {stot}')
src = {}
for key, n in sorted(byline.items(), key=lambda kv: (kv[0] or ('', 0))):
    if n / tot < 0.004 and samples[key] / max(stot, 1) < 0.004:
        continue
    print(f'{key[0] if key else "?":22s}:{key[1] if key else 0:4d}  instr {100*n/tot:5.2f}%  thr/instr {thr[key]/max(n,1):5.1f}  stall-samples {100*samples[key]/max(stot,1):5.2f}%')
