"""Turn gpurun_out ncu artefacts into the small text summaries committed under profiles/.
  python profiles/summarize.py launches gpurun_out/launches_X.csv > profiles/X_launches.txt
  python profiles/summarize.py full gpurun_out/prof_X.ncu-rep     > profiles/X_ncu_full.txt
"""
import collections
import csv
import subprocess
import sys


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    seq = []
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("p3tok::", "")
        agg.setdefault(name, []).append(v)
        seq.append((name, row.get("Grid Size", ""), row.get("Block Size", ""), v))
    tot = sum(sum(v) for v in agg.values())
    print(f"# {path}: {len(seq)} launches, {tot:.1f} us total (ncu per-launch times are cold-cache and serialised: compare shares)")
    print(f"{'kernel':45s} {'n':>4s} {'total_us':>10s} {'avg_us':>9s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:45s} {len(v):4d} {sum(v):10.1f} {sum(v) / len(v):9.1f} {100 * sum(v) / tot:6.1f}%")
    print("\n# launch sequence")
    for s in seq:
        print(f"{s[0]:45s} grid={s[1]:14s} block={s[2]:12s} {s[3]:9.1f} us")


WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ik = hdr.index("Kernel Name")
    print(f"# {path}: {len(data)} profiled launches (ncu --set full --clock-control none)")
    for r in data:
        print(f"\n## {r[ik].split('(')[0]}  grid={r[hdr.index('Grid Size')]} block={r[hdr.index('Block Size')]}")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:70s} {r[i]:>14s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
