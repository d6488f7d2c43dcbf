"""Turn gpurun_out ncu artefacts into the small text summaries committed under profiles/.
  python profiles/summarize.py launches gpurun_out/launches_X.csv > profiles/X_launches.txt
  python profiles/summarize.py full gpurun_out/prof_X.ncu-rep     > profiles/X_ncu_full.txt
  python profiles/summarize.py traffic gpurun_out/traffic_c2.csv c2:bf16:128 1   (merges into profiles/r02_traffic.json)
  python profiles/summarize.py sass adapting-.../p3tok/libp3tok.so           > profiles/r02_sass_opcodes.txt
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


EMBED_KERNELS = ("tc_", "rows_", "pad_weight", "partial_max", "group_max", "sgemm", "build_rows", "embed_", "apf_rel_rows", "fused_l1_pack")


def traffic(path, key, steps):
    """ncu csv with dram__bytes_read.sum / dram__bytes_write.sum (+ gpu__time_duration.sum) of `steps` eager steps
    (bench.py --ncu N under `ncu --profile-from-start off`): DRAM bytes of the patch-embedding launches per step ->
    profiles/r02_traffic.json[key], the number bench.py prints as roofline.traffic."""
    import json
    import os
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per_kernel = collections.OrderedDict()
    tot = {"embed": 0.0, "other": 0.0}
    us = {"embed": 0.0, "other": 0.0}
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("p3tok::", "")
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        m = row["Metric Name"]
        cls = "embed" if name.startswith(EMBED_KERNELS) else "other"
        if m.startswith("dram__bytes"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            tot[cls] += v
            per_kernel.setdefault(name, [0.0, 0.0, 0])
            per_kernel[name][0] += v
            if m.startswith("dram__bytes_read"):
                per_kernel[name][2] += 1
        elif m.startswith("gpu__time_duration"):
            v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
            us[cls] += v
            per_kernel.setdefault(name, [0.0, 0.0, 0])
            per_kernel[name][1] += v
    steps = int(steps)
    out_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "r02_traffic.json")
    d = json.load(open(out_path)) if os.path.isfile(out_path) else {}
    d[key] = {"embed_dram_bytes_per_step": tot["embed"] / steps, "other_dram_bytes_per_step": tot["other"] / steps,
              "embed_us_per_step_under_ncu": us["embed"] / steps,
              "per_kernel": {k: {"dram_bytes_per_step": v[0] / steps, "us_per_step_under_ncu": v[1] / steps, "launches_per_step": v[2] / steps}
                             for k, v in per_kernel.items()},
              "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum over the embed launches of {steps} eager step(s) ({os.path.basename(path)})"}
    json.dump(d, open(out_path, "w"), indent=1, sort_keys=True)
    print(f"{key}: embed {tot['embed'] / steps / 1e6:.1f} MB per step, other kernels {tot['other'] / steps / 1e6:.1f} MB")


SASS_OPS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UBLKCP", "STAS", "SYNCS",
            "CREDUX", "HMMA", "FFMA2", "FADD2", "FMUL2", "F2FP", "MUFU.EX2", "LDSM", "LDGSTS", "UCGABAR"]


def sass(lib):
    """Per-kernel SASS opcode counts of the built library (cuobjdump -sass): the tcgen05 / TMEM / TMA evidence."""
    import re
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            cur = cur.replace("void ", "").replace("p3tok::", "")
            counts.setdefault(cur, collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["_all"] += 1
            for want in SASS_OPS:
                if op == want or op.startswith(want + "."):
                    counts[cur][want] += 1
    print(f"# {lib}: SASS opcode counts per kernel (cuobjdump -sass, sm_100a)")
    print("# UTCHMMA = tcgen05.mma kind::f16 (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTMALDG/UTMASTG/UTMAREDG = TMA tensor load / store / "
          "reduce-add, UBLKCP = cp.async.bulk, STAS = st.async, CREDUX = redux.sync, HMMA = legacy mma.sync")
    used = [o for o in SASS_OPS if any(c[o] for c in counts.values())]
    print(f"{'kernel':70s} {'instrs':>7s} " + " ".join(f"{o:>8s}" for o in used))
    for k, c in counts.items():
        print(f"{k[:70]:70s} {c['_all']:7d} " + " ".join(f"{c[o]:8d}" for o in used))
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print(f"{'TOTAL':70s} {tot['_all']:7d} " + " ".join(f"{tot[o]:8d}" for o in used))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic, "sass": sass}[sys.argv[1]](*sys.argv[2:])
