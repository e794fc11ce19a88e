"""Static SASS opcode summary of the shipped library (no GPU needed): per kernel, instruction count and the
mnemonics that matter as evidence — UTMASTG / UTMALDG / UBLKCP (TMA), ATOMS / RED (shared atomics), LDG / STG widths,
MATCH / REDUX / VOTE (warp primitives), tensor-core mnemonics (none expected: there is no contraction on this path).
usage: python tools/sass_static.py [libflicb200.so] > profiles/rNN_sass_static.txt"""
import collections, os, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                        "fast-losless-image-compression-format_b200", "libflicb200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "--dump-resource-usage", so], capture_output=True, text=True).stdout
usage = {}
for m in re.finditer(r"Function (\S+):\n\s+REG:(\d+) STACK:(\d+) SHARED:(\d+)", res):
    usage[m.group(1)] = (int(m.group(2)), int(m.group(3)), int(m.group(4)))
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
cur, ops = None, collections.defaultdict(collections.Counter)
for ln in txt.split("\n"):
    m = re.match(r"\s+Function : (\S+)", ln)
    if m:
        cur = m.group(1); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
    if m and cur:
        ops[cur][m.group(2)] += 1
KEY = ("UTMASTG", "UTMALDG", "UBLKCP", "UTC", "HMMA", "LDTM", "STTM", "ATOMS", "RED", "ATOMG", "MATCH", "REDUX", "VOTE", "SHFL", "LDG", "STG", "LDS", "STS", "NANOSLEEP", "BAR", "CCTL", "PREFETCH")
print(f"# {os.path.basename(so)}: cuobjdump -sass, static instruction counts per kernel (sm_100a)")
for fn in sorted(ops, key=demangle):
    c = ops[fn]
    reg, stack, sh = usage.get(fn, (0, 0, 0))
    print(f"\n## {demangle(fn)}   [{sum(c.values())} SASS instructions, {reg} registers, {stack} B stack, {sh} B static smem]")
    sel = collections.Counter()
    for op, n in c.items():
        for k in KEY:
            if op.startswith(k):
                sel[".".join(op.split(".")[:3])] += n
    print("   " + ", ".join(f"{k} x{v}" for k, v in sorted(sel.items())))
tot = collections.Counter()
for c in ops.values():
    for op, n in c.items():
        tot[op.split(".")[0]] += n
print("\n# whole library:", ", ".join(f"{k} x{tot[k]}" for k in ("UTMASTG", "UTMALDG", "UBLKCP", "HMMA", "LDTM", "STTM", "ATOMS", "RED", "MATCH", "REDUX", "NANOSLEEP") if True))
