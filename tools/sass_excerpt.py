"""SASS excerpt of one loop of a kernel (evidence for the instruction counts quoted in DESIGN.md / profiles/README.md).
usage: python tools/sass_excerpt.py lib_or_obj <mangled-name substring> <loop index in tools/sass_loops.py order> > profiles/x_sass.txt"""
import collections, re, subprocess, sys
lib, pat, which = sys.argv[1], sys.argv[2], int(sys.argv[3])
names = [l.split()[-1] for l in subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines() if "Function :" in l]
name = [n for n in names if pat in n][0]
out = subprocess.run(["cuobjdump", "-sass", "-fun", name, lib], capture_output=True, text=True).stdout
ins = []
for l in out.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
idx = {a: i for i, (a, _) in enumerate(ins)}
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA.*0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) <= a and int(m.group(1), 16) in idx and i - idx[int(m.group(1), 16)] + 1 >= 20:
        loops.append((idx[int(m.group(1), 16)], i))
j, i = loops[which]
body = ins[j:i + 1]
mix = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in body)
fp64 = sum(v for k, v in mix.items() if k in ("DADD", "DMUL", "DFMA", "DSETP"))
packed = sum(v for k, v in mix.items() if k in ("FADD2", "FMUL2", "FFMA2"))
print(f"kernel {name}\nloop {body[0][0]:#x}..{body[-1][0]:#x}: {len(body)} instructions; FP64 {fp64}, packed fp32 {packed}, other {len(body) - fp64 - packed}")
print("mix: " + ", ".join(f"{k} {v}" for k, v in mix.most_common()))
print()
for a, t in body:
    print(f"{a:06x}  {t}")
