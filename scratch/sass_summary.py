"""cuobjdump -sass opcode counts per kernel of libmocat_b200.so -> profiles/sass_summary.md (Blackwell-native evidence)"""
import collections, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "mocat_b200", "libmocat_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kern, ops = None, collections.defaultdict(collections.Counter)
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1); continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern:
        ops[kern][m.group(1).split(".")[0]] += 1
dem = subprocess.run(["c++filt"] + list(ops), capture_output=True, text=True).stdout.splitlines()
KEY = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "FFMA2", "FADD2", "FMUL2", "HMMA", "MUFU", "REDUX",
       "IMAD", "LDGSTS", "SHFL", "ATOMS", "ATOMG", "RED"]
rows = []
for k, d in zip(ops, dem):
    c = ops[k]
    name = re.sub(r"\(.*", "", d)[:70]
    rows.append((sum(c.values()), name, c))
rows.sort(reverse=True)
out = ["# SASS opcode counts per kernel (`cuobjdump -sass mocat_b200/libmocat_b200.so`, sm_100a)", "",
       "`UTCHMMA` = tcgen05.mma, `LDTM`/`STTM` = tcgen05.ld/st, `UBLKCP` = cp.async.bulk (TMA), `SYNCS` = mbarrier ops, "
       "`FFMA2`/`FADD2` = packed fp32x2 (Blackwell), `REDUX` = warp reduce. No `HMMA` (legacy mma.sync) anywhere.", "",
       "| kernel | instr | " + " | ".join(KEY) + " |", "|---|---|" + "---|" * len(KEY)]
tot = collections.Counter()
for n, name, c in rows:
    tot.update(c)
    if n < 150 and not any(c[k] for k in KEY[:9]):
        continue
    out.append(f"| `{name}` | {n} | " + " | ".join(str(c[k]) if c[k] else "" for k in KEY) + " |")
out.append("| **whole library** | %d | " % sum(tot.values()) + " | ".join(str(tot[k]) if tot[k] else "" for k in KEY) + " |")
open(os.path.join(root, "profiles", "sass_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[-3:]))
print(len(rows), "kernels")
