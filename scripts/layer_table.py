"""profiles/rN_layers.md: the per-layer table of scripts/layer_times.py next to the pipe utilisation of an ncu --set full capture
of launches 13..27 of a forward (L6 conv, L6 filter, ..., L13 conv; scripts/r2_final.sh): layer_table.py layers.txt rep.ncu-rep"""
import csv, io, re, subprocess, sys
layers_txt, rep = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
col = lambda name: hdr.index(name)
M = {"ms": "gpu__time_duration.sum", "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
     "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
     "regs": "launch__registers_per_thread", "xbar": "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed"}
units = rows[1]
def gb(r, key):
    v, u = float(r[col(M[key])]), units[col(M[key])]
    return v * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}[u]
def ms(r):
    v, u = float(r[col(M["ms"])]), units[col(M["ms"])]
    return v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1.0)
cap = {}
# capture order (--launch-skip 41 --launch-count 15 over the conv / filter kernels): L5 filter, L6 conv, L6 filter, L7 conv, ..., L12 filter
first_is_filter = "flrelu" in rows[2][col("Kernel Name")]
for i, r in enumerate(rows[2:]):
    j = i + 1 if first_is_filter else i + 2      # j even: conv of layer 5 + j / 2, j odd: filter of layer 5 + (j - 1) / 2
    layer, kind = (5 + j // 2, "conv") if j % 2 == 0 else (5 + (j - 1) // 2, "flrelu")
    name = re.sub(r"\(.*", "", r[col("Kernel Name")]).replace("void ", "").replace("unnamed>::", "")
    cap[(layer, kind)] = (name, ms(r), float(r[col(M["tensor"])]), float(r[col(M["issue"])]), gb(r, "rd") + gb(r, "wr"), r[col(M["regs"])], float(r[col(M["xbar"])]))
print("# Per-layer table, StyleGAN3-T 1024^2, 16 frames per step (one B200)\n")
print("`ms` / `TFLOP/s` / `TB/s`: device time inside back-to-back forwards (CUDA events around every launch, scripts/layer_times.py; power-capped clocks).")
print("`ncu` columns: the same launch isolated under `ncu --set full --clock-control none` (scripts/r2_final.sh; the launches from the L5 filter to the L12 filter): duration, tensor pipe busy")
print("(tcgen05 for the conv, HMMA for the filter), issue slots busy, DRAM bytes read + written, L1TEX->XBAR request port busy.\n")
print("| layer | conv kernel | conv ms | useful TFLOP/s | ncu ms | tensor pipe % | issue % | DRAM GB | xbar % | filter ms | alg. TB/s | ncu ms | HMMA pipe % | issue % | DRAM GB |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for line in open(layers_txt):
    m = re.match(r"(L(\d+)_\S+)\s+conv\s+([\d.]+) ms \(\s*(\d+) TFLOP.*flrelu\s+([\d.]+) ms \(\s*([\d.]+) TB", line)
    if not m or m.group(3) == "0.000":
        continue
    L = int(m.group(2))
    c, f = cap.get((L, "conv")), cap.get((L, "flrelu"))
    cc = f"`{c[0]}` | {m.group(3)} | {m.group(4)} | {c[1]:.3f} | {c[2]:.1f} | {c[3]:.1f} | {c[4]:.2f} | {c[6]:.0f}" if c else f" | {m.group(3)} | {m.group(4)} | | | | | "
    ff = f"{m.group(5)} | {m.group(6)} | {f[1]:.3f} | {f[2]:.1f} | {f[3]:.1f} | {f[4]:.2f}" if f else f"{m.group(5)} | {m.group(6)} | | | | "
    print(f"| {m.group(1)} | {cc} | {ff} |")
for line in open(layers_txt):
    if line.startswith("B=") or line.startswith("{"):
        print("\n" + line.strip())
