"""Summarise an `ncu --page raw --csv` export of igemm_kernel launches (one warm forward):
per-launch duration, DRAM bytes, tensor-pipe activity; writes the text table and the per-launch
average DRAM traffic bench.py reports as roofline.traffic.
usage: python scripts/summarize_ncu.py gpurun_out/r1b_igemm_raw.csv profiles/r1b_igemm_ncu_summary.txt profiles/igemm_traffic.json"""
import csv, json, re, sys

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
src, out_txt, out_json = sys.argv[1:4]
rows = list(csv.reader(open(src)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {n: i for i, n in enumerate(hdr)}


def val(r, name):
    try:
        return float(r[ix[name]]) * SCALE.get(units[ix[name]], 1.0)
    except ValueError:
        return float("nan")


lines = ["# ncu --set full --clock-control none, igemm_kernel launches of one warm configs[1] forward (B=256, V=2, bf16)",
         "# us = gpu__time_duration.sum (cold-cache, serialised); tensor% = sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed;",
         "# dram% = dram__throughput.avg.pct_of_peak_sustained_elapsed; lts% = lts__throughput.avg.pct_of_peak_sustained_elapsed",
         f"{'id':>3s} {'kernel / <N,RES,F32,HALO,STATS>':>16s} {'grid':>5s} {'us':>8s} {'rd MB':>8s} {'wr MB':>8s} {'GB/s':>7s} {'tensor%':>8s} {'dram%':>6s} {'lts%':>6s}"]
tot_t = tot_b = 0.0
for r in data:
    name = r[ix["Kernel Name"]]
    m = re.search(r"igemm_kernel<([^>]*)>", name)
    if m:
        cfg = m.group(1).replace("(int)", "").replace("(bool)", "").replace(" ", "")
    else:
        m2 = re.search(r"(\w+_kernel)", name)
        cfg = (m2.group(1) if m2 else name)[:16]
    t = val(r, "gpu__time_duration.sum")
    rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    tp = val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")
    dp = val(r, "FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed")
    lp = val(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed")
    grid = r[ix["Grid Size"]].strip("()").split(",")[0]
    tot_t += t; tot_b += rd + wr
    lines.append(f"{r[ix['ID']]:>3s} {cfg:>16s} {grid:>5s} {t:8.1f} {rd / 1e6:8.1f} {wr / 1e6:8.1f} {(rd + wr) / t / 1e3:7.0f} {tp:8.1f} {dp:6.1f} {lp:6.1f}")
n = len(data)
lines.append(f"# {n} launches: total {tot_t:.0f} us, DRAM traffic {tot_b / 1e9:.3f} GB -> {tot_b / n / 1e6:.1f} MB per launch, {tot_b / tot_t / 1e3:.0f} GB/s average")
open(out_txt, "w").write("\n".join(lines) + "\n")
json.dump({"kernel": "igemm_kernel", "launches": n, "dram_bytes_per_launch": tot_b / n, "dram_bytes_total": tot_b,
           "source": src, "how": "ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum, mean over the launches of one forward"},
          open(out_json, "w"), indent=1)
print(lines[-1])
