"""Per-SASS-instruction stall samples from an `ncu --page source --print-source cuda,sass --csv` export
(deduplicated by address), grouped by opcode class, plus the top instructions.
usage: sass_hotspots.py FILE.csv.gz [TOP]"""
import csv, gzip, sys, collections
rows = csv.reader(gzip.open(sys.argv[1], 'rt'))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
seen = {}
cur_line = None
for r in rows:
    if not r: continue
    if r[0] in ('File Path', 'Function Name', 'Line No'):
        if r[0] == 'File Path': cur_file = r[1].split('/')[-1]
        continue
    if r[0].isdigit():
        cur_line = (cur_file, int(r[0]), r[1][:70]); continue
    if r[0] == '' and len(r) >= 8 and r[2].startswith('0x'):
        addr = int(r[2], 16)
        try: s, ni, ex = int(r[4]), int(r[5]), int(r[7])
        except ValueError: continue
        if addr not in seen: seen[addr] = [r[3].strip(), s, ni, ex, cur_line]
tot = sum(v[1] for v in seen.values())
print("instructions", len(seen), "samples", tot)
cls = collections.Counter(); cnt = collections.Counter(); exe = collections.Counter()
for a, (txt, s, ni, ex, ln) in seen.items():
    op = txt.split()[1] if txt.startswith('@') else txt.split()[0]
    op = op.split('.')[0]
    cls[op] += s; cnt[op] += 1; exe[op] += ex
print("--- by opcode: samples%  static  executed(M)")
for op, s in cls.most_common(18):
    print(f"{op:10s} {100*s/tot:5.1f}%  {cnt[op]:5d}  {exe[op]/1e6:9.1f}")
print("--- top instructions")
for a, (txt, s, ni, ex, ln) in sorted(seen.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100*s/tot:5.2f}%  {txt[:60]:60s} {ln[0]}:{ln[1]}")
# per-file / per-line-range phase table: sass_hotspots.py FILE TOP file:lo-hi=name,...
if len(sys.argv) > 3:
    print("--- phases (samples%, executed M warp-instr)")
    rest_s, rest_e = tot, sum(v[3] for v in seen.values())
    for spec in sys.argv[3].split(','):
        rng, name = spec.split('=')
        f, lh = rng.split(':'); lo, hi = map(int, lh.split('-'))
        s = sum(v[1] for v in seen.values() if v[4] and v[4][0] == f and lo <= v[4][1] <= hi)
        e = sum(v[3] for v in seen.values() if v[4] and v[4][0] == f and lo <= v[4][1] <= hi)
        rest_s -= s; rest_e -= e
        print(f"{name:14s} {100*s/tot:5.1f}%  {e/1e6:9.1f}")
    print(f"{'other':14s} {100*rest_s/tot:5.1f}%  {rest_e/1e6:9.1f}")
