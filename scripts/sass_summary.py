"""SASS evidence per kernel of libb200forge.so (cuobjdump -sass): counts of the mnemonics that prove the Blackwell-native path
(B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA tensor loads, FFMA2/FADD2 = packed fp32,
ACQBULK / PREEXIT = griddepcontrol (programmatic dependent launch).
Usage: python scripts/sass_summary.py > profiles/sass_r2_summary.txt"""
import collections
import os
import re
import subprocess

so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "stable-diffusion-webui-forge_b200", "libb200forge.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()  # noqa: E731
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "SYNCS", "MUFU.EX2", "FFMA2", "FADD2", "HMMA", "ACQBULK", "PREEXIT"]
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        c = counts[cur]
        c["_n"] += 1
        for k in KEYS:
            if k != "UTCHMMA.2CTA" and (op == k or op.startswith(k + ".")):
                c[k] += 1
        if op.startswith("UTCHMMA") and ".2CTA" in op:
            c["UTCHMMA.2CTA"] += 1
print(f"# cuobjdump -sass {os.path.basename(so)} — instruction counts per kernel (static, per compiled function)")
print("# " + " ".join(f"{k:>12s}" for k in ["instrs"] + KEYS) + "  kernel")
tot = collections.Counter()
for fn, c in counts.items():
    name = re.sub(r"\(.*", "", demangle(fn)).replace("b200::", "")
    print("  " + " ".join(f"{c.get(k, 0):12d}" for k in ["_n"] + KEYS) + "  " + name)
    tot.update(c)
print("  " + " ".join(f"{tot.get(k, 0):12d}" for k in ["_n"] + KEYS) + "  TOTAL")
