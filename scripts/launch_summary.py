"""profiles/*.md from an ncu launch list (gpu__time_duration.sum CSV): launch_summary.py in.csv out.md "title" """
import collections, csv, re, sys
src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(src, errors="ignore")))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"^void ", "", r[ki]).replace("mb::<unnamed>::", "").replace("mb::", "")
    name = re.sub(r"\((?!anonymous).*$", "", name)
    v = float(r[vi].replace(",", ""))
    ms = v / 1e6 if r[ui].startswith("ns") else (v / 1e3 if r[ui].startswith("us") else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
tot, n = sum(a[1] for a in agg.values()), sum(a[0] for a in agg.values())
lines = [f"# {title}", "", "(first 400 launches of the command: set-up kernels, the audio feature pass, warm-up and timed forwards; cold-cache and "
         "serialised under ncu: compare SHARES with bench.py's `kernel_ms_per_step`, not absolutes)", "",
         "| kernel | launches | total ms | share |", "|---|---|---|---|"]
for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if ms / tot >= 0.001:
        lines.append(f"| {k} | {c} | {ms:.3f} | {100 * ms / tot:.1f}% |")
lines.append(f"| total | {n} | {tot:.3f} | 100% |")
fam = lambda pre: sum(ms for k, (c, ms) in agg.items() if k.startswith(pre))
conv, fl = fam("conv_"), fam("flrelu")
rest = sum(ms for k, (c, ms) in agg.items() if k.split("<")[0] in ("input_feat_kernel", "input_prep_kernel", "styles_kernel", "torgb_out_kernel"))
synth = conv + fl + rest
lines += ["", f"Synthesis kernels only ({synth:.1f} ms): modulated conv {100 * conv / synth:.1f} %, filtered_lrelu {100 * fl / synth:.1f} %, "
          f"input / styles / ToRGB {100 * rest / synth:.1f} %."]
open(dst, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[-3:]))
