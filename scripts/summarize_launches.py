"""Summarise an `ncu --metrics ... --csv --log-file` launch list (long format: one row per launch
and metric): per kernel name -> launches, total time, share of the step, DRAM bytes, GB/s, mean
tensor-pipe activity. usage: summarize_launches.py <csv> <out.txt> [traffic.json kernel-regex]"""
import collections, csv, json, re, sys

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "%": 1.0}
src, out = sys.argv[1:3]
rows = list(csv.reader(open(src, errors="replace")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]; ix = {n: i for i, n in enumerate(hdr)}
launch = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    d = launch.setdefault(r[ix["ID"]], {"name": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]]})
    try:
        d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", "")) * SCALE.get(r[ix["Metric Unit"]], 1.0)
    except ValueError:
        pass


def short(name):
    m = re.search(r"(\w+)(<[^(]*>)?\(", name)
    base = m.group(1) if m else name[:40]
    t = re.search(r"igemm_kernel<([^>]*)>", name)
    if t:
        base += "<" + t.group(1).replace("(int)", "").replace("(bool)", "").replace(" ", "") + ">"
    t = re.search(r"(bn_\w+|wgrad_kernel|maxpool\w*|relu_bwd_kernel|colsum_kernel|avgpool\w*|rotate_gather_kernel|head_loss\w*|tstat_kernel|splitk_reduce_kernel|mask_bits_kernel)<([^>]*)>", name)
    if t and "igemm" not in base:
        base = t.group(1) + "<" + t.group(2).replace("__nv_bfloat16", "bf16").replace("(int)", "").replace("(bool)", "").replace(" ", "") + ">"
    return base


agg = collections.OrderedDict()
tot = 0.0
for d in launch.values():
    k = short(d["name"])
    a = agg.setdefault(k, [0, 0.0, 0.0, 0.0])
    t = d.get("gpu__time_duration.sum", 0.0)
    a[0] += 1; a[1] += t; tot += t
    a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    a[3] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) * t
lines = [f"# {src}: {len(launch)} launches, {tot:.0f} us in total (ncu --clock-control none, per-launch times are",
         "# cold-cache and serialised: compare shares). GB/s = (dram read + write bytes) / time; tensor% = time-weighted",
         "# sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed.",
         f"{'kernel':58s} {'n':>4s} {'us':>9s} {'share':>6s} {'dram MB':>9s} {'GB/s':>6s} {'tensor%':>7s}"]
for k, (n, t, b, tp) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{k:58s} {n:4d} {t:9.1f} {100 * t / tot:5.1f}% {b / 1e6:9.1f} {b / t / 1e3 if t else 0:6.0f} {tp / t if t else 0:7.1f}")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:24]))
if len(sys.argv) > 4:
    pat = re.compile(sys.argv[4])
    sel = [d for d in launch.values() if pat.search(d["name"])]
    b = sum(d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0) for d in sel)
    json.dump({"kernel": sys.argv[4], "launches": len(sel), "dram_bytes_per_launch": b / max(len(sel), 1), "dram_bytes_total": b,
               "source": src, "how": "ncu launch list of one warm forward: dram__bytes_read.sum + dram__bytes_write.sum, mean over the kernel's launches"},
              open(sys.argv[3], "w"), indent=1)
