"""Static SASS census of libmaua_b200.so: per kernel, counts of the mnemonics that prove which hardware path it uses
(UTCHMMA / UTCQMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit,
HMMA = legacy mma.sync, LDGSTS = cp.async, SYNCS = mbarrier).  Usage: sass_census.py [lib.so] > profiles/rN_sass_census.md"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "maua_b200", "lib", "libmaua_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", txt)), capture_output=True, text=True).stdout.split("\n")
KEYS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "HMMA", "LDGSTS", "SYNCS", "LDSM", "STG", "LDG"]
cur, rows, order = None, collections.defaultdict(collections.Counter), []
fi = 0
for line in txt.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = re.sub(r"\(.*", "", names[fi].replace("(anonymous namespace)::", "")).replace("void ", "") + f" #{fi}"; fi += 1; order.append(cur); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        rows[cur]["_n"] += 1
        for k in KEYS:
            if m.group(1) == k or (k in ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "HMMA", "LDGSTS", "SYNCS", "LDSM") and m.group(1).startswith(k)):
                rows[cur][k] += 1
print("# Static SASS census of libmaua_b200.so (cuobjdump -sass, sm_100a): instruction counts per kernel\n")
print("| kernel | instrs | " + " | ".join(KEYS) + " |"); print("|---|---|" + "---|" * len(KEYS))
for n in order:
    c = rows[n]
    print(f"| `{n[:70]}` | {c['_n']} | " + " | ".join(str(c[k]) if c[k] else "" for k in KEYS) + " |")
