"""Per-source-line instruction counts of one kernel: joins `ncu --page source --print-source sass --csv`
(per-SASS-instruction executed counts) with `nvdisasm -g -c` line annotations of the same cubin.
usage: python tools/ncu_lines.py sass.csv disasm.txt MANGLED_NAME_SUBSTR [top]"""
import csv, re, sys
sass_csv, disasm, sym = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# address -> (file line) from nvdisasm
line_of, cur, inside = {}, None, False
for ln in open(disasm, errors="replace"):
    if ln.startswith("\t.section\t.text."):
        inside = sym in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ia, ie, it, iss = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
base = None
tot, per, samp = 0, {}, {}
for r in rows[hi + 1:]:
    if len(r) <= ie or not r[ie]:
        continue
    a = int(r[ia], 16)
    base = a if base is None else base
    n = int(r[ie]); s = int(r[iss] or 0)
    key = line_of.get(a - base, ("?", 0))
    per[key] = per.get(key, 0) + n
    samp[key] = samp.get(key, 0) + s
    tot += n
print("total warp instructions", tot, "samples", sum(samp.values()))
src = {}
for (f, l), n in sorted(per.items(), key=lambda kv: -kv[1])[:top]:
    if f not in src:
        try: src[f] = open("/root/repo/fast-losless-image-compression-format_b200/csrc/" + f).read().split("\n")
        except Exception: src[f] = []
    text = src[f][l - 1].strip()[:90] if 0 < l <= len(src[f]) else ""
    print(f"{100*n/tot:5.1f}%  samp {100*samp[(f,l)]/max(1,sum(samp.values())):5.1f}%  {f}:{l}  {text}")
