"""Static SASS opcode histogram of the shipped library, per kernel (cuobjdump -sass): which tensor / async-copy /
Blackwell-specific instructions each kernel contains.  CPU only (cross-compiled objects).
Usage: python tools/sass_histogram.py [klhr_b200/libklhr_sm100.so] > profiles/rNN_sass_opcodes.txt"""
import collections, re, subprocess, sys

lib = sys.argv[1] if len(sys.argv) > 1 else "klhr_b200/libklhr_sm100.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, per = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(klhr::StepArgs\)|klhr::", "", kern)
        per[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", line)
    if m and kern:
        full = m.group(1)
        base = full.split(".")[0]
        name = ".".join(full.split(".")[:2]) if base in ("DMMA", "LDGSTS", "MUFU", "UBLKCP", "UTMALDG", "SYNCS", "LDS", "STS", "LDG", "STG") else base
        per[kern][name] += 1
watch = ("DMMA", "HMMA", "UTC", "LDTM", "STTM", "UTMA", "UBLKCP", "LDGSTS", "LDGDEPBAR", "SYNCS", "DFMA", "DADD", "DMUL", "MUFU", "IMAD.WIDE", "LDL", "STL")
print(f"# cuobjdump -sass {lib}: static instruction counts per kernel (entry functions with >= 400 instructions)")
print("# tcgen05 (UTC*MMA), TMEM (LDTM/STTM) and TMA (UTMA*/UBLKCP) do not appear: every number on this path is fp64, which")
print("# tcgen05 / TMEM have no data type for (DESIGN.md section 4); the tensor-core instruction that applies is DMMA, the")
print("# asynchronous copy that applies is LDGSTS (cp.async), both in the dense kernel and the pooled-PCA SYRK")
for k, c in per.items():
    tot = sum(c.values())
    if tot < 400:
        continue
    sel = {w: sum(v for n, v in c.items() if n.startswith(w)) for w in watch}
    print(f"{tot:7d}  {k[:110]}")
    print("         " + "  ".join(f"{w} {v}" for w, v in sel.items() if v))
