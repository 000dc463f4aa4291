"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export per CUDA source line.
usage: src_hotspots.py FILE.csv.gz [TOP]"""
import csv, gzip, sys, collections
rows = list(csv.reader(gzip.open(sys.argv[1], 'rt')))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
line = None; src = {}
samples = collections.Counter(); insts = collections.Counter()
fn = ''; fpath = ''
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fpath = r[1].rsplit('/', 1)[-1]; continue
    if r[0] == 'Function Name': fn = fpath; continue          # key on the FILE: line numbers of different headers collide
    if r[0] == 'Line No': continue
    if r[0].isdigit():
        line = (fn, int(r[0])); src[line] = r[1][:90]
        continue
    if r[0] == '' and len(r) >= 8 and r[2].startswith('0x') and line is not None:
        try:
            samples[line] += int(r[6]); insts[line] += int(r[7])
        except ValueError:
            pass
tot = sum(samples.values())
print("total samples", tot)
for ln, s in samples.most_common(top):
    print(f"{100*s/tot:5.1f}%  {insts[ln]/1e6:9.1f}M  {ln[0][:28]:28s} L{ln[1]:4d}  {src[ln]}")

print("--- per file")
byfile = collections.Counter(); ibyfile = collections.Counter()
for (f, ln), v in samples.items(): byfile[f] += v; ibyfile[f] += insts[(f, ln)]
for f, v in byfile.most_common(): print(f"{100*v/tot:5.1f}%  {ibyfile[f]/1e6:9.1f}M  {f}")

# optional phase table: src_hotspots.py FILE TOP name:lo-hi,name:lo-hi,...
if len(sys.argv) > 3:
    print("--- phases")
    for spec in sys.argv[3].split(','):
        name, rng = spec.split(':'); lo, hi = map(int, rng.split('-'))      # name may be file.cuh@label
        fsel = name.split('@')[0] if '@' in name else None
        s = sum(v for (f, ln), v in samples.items() if lo <= ln <= hi and (fsel is None or f == fsel))
        print(f"{name:14s} {100*s/tot:5.1f}%")
