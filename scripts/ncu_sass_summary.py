"""Summarise an `ncu -i X.ncu-rep --page source --csv` export: stall samples by opcode and the hottest instructions."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        sections.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
sec = sections[which]
hdr, data = sec["hdr"], sec["data"]
ix = {k: i for i, k in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]]) for r in data)
print(sec["name"][:100], "| samples", tot, "| static instr", len(data))
byop, cnt, execd = collections.Counter(), collections.Counter(), collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
    op = m.group(2).split(".")[0] if m else "?"
    byop[op] += int(r[ix["# Samples"]]); cnt[op] += 1; execd[op] += int(r[ix["Instructions Executed"]])
te = sum(execd.values())
for op, s in byop.most_common(22):
    print(f"{op:12s} samples {s:7d} {100*s/tot:5.1f}%  static {cnt[op]:5d}  exec {execd[op]:10d} {100*execd[op]/te:5.1f}%")
print()
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 20]:
    st = {k[6:]: int(r[ix[k]]) for k in hdr if k.startswith("stall_") and "Not Issued" not in k and r[ix[k]] not in ("0", "")}
    print(r[ix["Source"]][:64].ljust(64), r[ix["# Samples"]].rjust(6), sorted(st.items(), key=lambda kv: -kv[1])[:3])
