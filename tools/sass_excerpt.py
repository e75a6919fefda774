"""SASS evidence for DESIGN.md's instruction-level claims:  python tools/sass_excerpt.py <object or .so> <mangled-substring> out.md "title"

Writes the opcode histogram of the kernel and the hot loop from the tile loads (LDG.E.*.128) to the loop's back edge."""
import collections
import re
import subprocess
import sys

obj, key, out, title = sys.argv[1:5]
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, funcs = None, {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if cur and m:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
name = [k for k in funcs if key in k][0]
ins = funcs[name]
dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
op = lambda s: re.sub(r"^@!?U?P\d+\s+", "", s).split()[0]
hist = collections.Counter(op(s) for _, s in ins)
# the hot loop = the innermost loop around the full-tile loads: the backward branch with the SHORTEST span that
# encloses four consecutive 128-bit loads (2 x-vectors + 2 v-vectors; 2 in a 24-byte pass ... still >= 2 pairs)
addr_index = {a: i for i, (a, _) in enumerate(ins)}
ldg = [i for i, (_, s) in enumerate(ins) if "LDG.E" in s and ".128" in s and ".NA" in s]
if len(ldg) < 4:                                        # not a texture-gather kernel: the evict-first loads
    ldg = [i for i, (_, s) in enumerate(ins) if "LDG.E" in s and ".128" in s]
best = None
for i, (a, s) in enumerate(ins):
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", s)
    if not m:
        continue
    t = int(m.group(1), 16)
    if t >= a or t not in addr_index:
        continue
    j = addr_index[t]
    inside = [k for k in ldg if j <= k <= i]
    if len(inside) >= 4 and (best is None or i - j < best[1] - best[0]):
        best = (j, i)
loop = ins[best[0]:best[1] + 1]
lh = collections.Counter(op(s).split(".")[0] for _, s in loop)
with open(out, "w") as f:
    f.write("# %s\n\n`%s`\n\nfrom `cuobjdump -sass` of the shipped library (sm_100a).  %d instructions in the kernel; hot loop "
            "(one tile = 4 particles per thread, including the cold fall-back blocks laid out inside it) = %d instructions.\n\n"
            % (title, dem, len(ins), len(loop)))
    f.write("## What to look for\n\n")
    checks = [("LDG.E.EF.128 / STG.E.EF.128", "16-byte evict-first particle loads / stores", ["LDG.E.EF.128", "STG.E.EF.128"]),
              ("LDG.E.NA.128", "texture-gather kernels: 16-byte particle loads that do not allocate in the L1 (it holds the field table)", ["LDG.E.NA.128"]),
              ("TLD.LZ", "texture-gather kernels: the (E_j, E_j+1) gather as a texture fetch (texture pipe, not the LSU data pipe)", ["TLD"]),
              ("UBLKPF.L2", "bulk L2 prefetch of the tile two iterations ahead (one thread)", ["UBLKPF.L2"]),
              ("DFMA.RM", "floor(x a) by a round-down FMA against the magic constant (cell index bracket, 2 per position)", ["DFMA.RM"]),
              ("ATOMS.ADD / ATOMS.POPC.INC", "native 32-bit shared atomics of the split deposit; no ATOMS.CAST.SPIN (CAS loop)", ["ATOMS.ADD", "ATOMS.POPC.INC.32", "ATOMS.CAST.SPIN"]),
              ("LDS.128", "one 16-byte shared load per gather (E_j, E_j+1)", ["LDS.128"]),
              ("LDL / STL", "local-memory (spill) traffic", ["LDL", "STL"]),
              ("F2I / I2F / FRND on fp64", "conversion-pipe instructions (avoided: floor and fixed-point conversion run on the fp64 add pipe)", ["F2I", "I2F", "FRND"])]
    f.write("| instruction | meaning | in kernel | in hot loop |\n|---|---|---|---|\n")
    for label, meaning, pats in checks:
        ck = sum(1 for _, s in ins if any(op(s).startswith(p) for p in pats))
        cl = sum(1 for _, s in loop if any(op(s).startswith(p) for p in pats))
        f.write("| `%s` | %s | %d | %d |\n" % (label, meaning, ck, cl))
    f.write("\n## Opcode histogram of the hot loop (static)\n\n```\n")
    for k, v in lh.most_common(30):
        f.write("%-12s %d\n" % (k, v))
    f.write("```\n\n## Hot loop, first particle of the tile (loads, stage-2 drift redone, cell bracket, gather, kick, drift, "
            "deposit; state wrap + next stage-0 deposit in the final pass)\n\n```\n")
    n_atoms, shown = 0, 0
    for a, s in loop:
        f.write("/*%04x*/ %s\n" % (a, s))
        shown += 1
        if "ATOMS" in s:
            n_atoms += 1
        if shown > 420 or n_atoms >= 12:
            break
    f.write("...\n```\n")
print(open(out).read()[:3000])
