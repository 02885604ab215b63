"""Join an ncu source-page CSV (per-SASS-address executed counts / stall samples) with `nvdisasm -g` line info.

usage: line_profile.py <ncu --page source --csv file> <nvdisasm -g -c output> <mangled kernel substring> [topN]
Prints executed warp-instructions and stall samples per source line of the kernel (all inlined code attributed
to the innermost source line nvdisasm reports)."""
import collections
import csv
import re
import sys

src_csv, dis, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(src_csv)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ia, ist = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
recs = []
for r in rows[h + 1:]:
    if len(r) <= ia:
        continue
    try:
        recs.append((int(r[0], 16) if r[0].startswith("0x") else int(r[0]), int(r[ia]), int(r[ist] or 0)))
    except ValueError:
        pass
base = min(a for a, _, _ in recs)
byoff = {a - base: (n, s) for a, n, s in recs}
line_of = {}
cur, infn = None, False
for l in open(dis):
    if l.startswith(".text."):
        infn = kname in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m:
        line_of[int(m.group(1), 16)] = cur
ex, st, cnt = collections.Counter(), collections.Counter(), collections.Counter()
for off, (n, s) in byoff.items():
    k = line_of.get(off, ("?", 0))
    ex[k] += n
    st[k] += s
    cnt[k] += 1
tot, tots = sum(ex.values()), sum(st.values())
print(f"static SASS instructions {len(byoff)}, executed warp-instructions {tot}, stall samples {tots}")
for k, v in ex.most_common(top):
    print(f"{k[0]}:{k[1]:<5d} exec {v:9d} {100 * v / tot:5.1f}%  stall {100 * st[k] / max(tots, 1):5.1f}%  static {cnt[k]}")
