"""List the loops (backward branches) of one kernel in a cubin/.so with their static instruction mix.
usage: python tools/sass_loops.py lib.so <substring of mangled kernel name> [min_len]"""
import re, subprocess, sys, collections
lib, pat = sys.argv[1], sys.argv[2]
minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 20
names = [l.split()[-1] for l in subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines() if "Function :" in l]
names = [n for n in names if pat in n]
for name in names:
    out = subprocess.run(["cuobjdump", "-sass", "-fun", name, lib], capture_output=True, text=True).stdout
    ins = []
    for l in out.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m: ins.append((int(m.group(1), 16), m.group(2)))
    print(f"== {name}: {len(ins)} instructions")
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?`?\(?\.?L?_?x?_?\d*\)?\s*0x([0-9a-f]+)", t) or re.search(r"BRA.*0x([0-9a-f]+)", t)
        if not m or "BRA" not in t: continue
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr_index:
            j = addr_index[tgt]
            if i - j + 1 < minlen: continue
            mix = collections.Counter()
            for _, tt in ins[j:i + 1]:
                tt = re.sub(r"^@!?U?P\d+\s+", "", tt)
                mix[tt.split()[0].split(".")[0]] += 1
            n = i - j + 1
            dp = sum(v for k, v in mix.items() if k in ("DADD", "DMUL", "DFMA", "DSETP", "DMNMX"))
            print(f"  loop {tgt:#x}..{a:#x}: {n} instrs, FP64 {dp}, other {n - dp}, port-model cycles {2 * dp + n - dp}: " + ", ".join(f"{k} {v}" for k, v in mix.most_common(14)))
