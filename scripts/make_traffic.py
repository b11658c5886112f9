"""profiles/traffic.json from an ncu CSV of one forward (merged per batch size: keys ..._b8, ..._b16):
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      -k regex:'flrelu|conv_' --launch-skip 28 --launch-count 28 --csv --log-file gpurun_out/traffic.csv \
      python scripts/one_forward.py 8 T        # then: make_traffic.py gpurun_out/traffic.csv profiles/traffic.json 8
bench.py reports these measured DRAM bytes per step as roofline.traffic next to the algorithmic bytes."""
import csv, json, sys
src, dst, B = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 8
tag = (sys.argv[4] + "_") if len(sys.argv) > 4 and sys.argv[4] != "T" else ""   # "R": keys ..._R_b16 (StyleGAN3-R)
rows = [r for r in csv.reader(open(src)) if len(r) >= 15 and r[0].isdigit()]
acc = {"filtered_lrelu": {"bytes": 0.0, "ns": 0.0, "launches": set()}, "modulated_conv2d": {"bytes": 0.0, "ns": 0.0, "launches": set()}}
for r in rows:
    name, metric, unit, val = r[4], r[12], r[13], float(r[14].replace(",", ""))
    fam = "filtered_lrelu" if "flrelu" in name else ("modulated_conv2d" if "conv_" in name else None)
    if fam is None:
        continue
    acc[fam]["launches"].add(r[0])
    if metric.startswith("dram__bytes"):
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        acc[fam]["bytes"] += val * scale
    elif metric.startswith("gpu__time_duration"):
        acc[fam]["ns"] += val * {"ns": 1.0, "us": 1e3, "ms": 1e6}[unit]
import os
out = json.load(open(dst)) if os.path.exists(dst) else {}
out["source"] = "ncu dram__bytes_read.sum + dram__bytes_write.sum over the launches of ONE forward per batch size, scripts/make_traffic.py"
for fam, d in acc.items():
    out["%s_bytes_per_step_%sb%d" % (fam, tag, B)] = d["bytes"]
    out["%s_launches" % fam] = len(d["launches"])
    out["%s_ncu_ms_%sb%d" % (fam, tag, B)] = d["ns"] / 1e6
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out))
