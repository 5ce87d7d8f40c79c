"""Turn the ncu CSV exports under gpurun_out/ into the tracked evidence under profiles/:
   profiles/<round>_ncu_<tag>_details.csv   (ncu --page details, as exported)
   profiles/<round>_ncu_key_metrics.json    (selected raw metrics per capture)
   profiles/ncu_summary.json                (what bench.py reads: DRAM bytes per launch of the dominant kernel)
   profiles/<round>_launches_bench.csv + a per-kernel share table
usage: summarize_profiles.py r01"""
import collections
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"

KEYS = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum")


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    os.makedirs(P, exist_ok=True)
    key_metrics, summary = {}, {}
    for tag in ("ccsaq_m4", "mma_m4"):
        raw = os.path.join(G, f"prof_{tag}_raw.csv")
        if not os.path.exists(raw):
            continue
        rows = list(csv.reader(open(raw)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        key_metrics[tag] = {"kernel": d.get("Kernel Name", ("", ""))[0]}
        for k in KEYS:
            if k in d:
                key_metrics[tag][k] = f"{d[k][0]} {d[k][1]}".strip()
        rd = to_bytes(*d["dram__bytes_read.sum"])
        wr = to_bytes(*d["dram__bytes_write.sum"])
        dur = float(d["gpu__time_duration.sum"][0].replace(",", ""))
        dur_us = dur / 1e3 if d["gpu__time_duration.sum"][1] == "ns" else dur
        summary[tag] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr, "duration_us_under_ncu": dur_us,
                        "algorithmic_bytes": 8 * 10_000_000 * 9}
        shutil.copy(os.path.join(G, f"prof_{tag}_details.csv"), os.path.join(P, f"{rnd}_ncu_{tag}_details.csv"))
        # instruction mix + stall reasons from the source page
        src = os.path.join(G, f"prof_{tag}_source.csv")
        if os.path.exists(src):
            rows = list(csv.reader(open(src)))
            h2, data = rows[1], rows[2:]
            iS, iI = h2.index("Source"), h2.index("Instructions Executed")
            ops, st = collections.Counter(), collections.Counter()
            stall_cols = [(i, h) for i, h in enumerate(h2) if h.startswith("stall_") and "Not Issued" not in h]
            for r in data:
                if not r[iI].isdigit():
                    continue
                toks = r[iS].strip().split()
                op = toks[1] if toks[0].startswith("@") else toks[0]
                ops[op.split(".")[0]] += int(r[iI])
                for i, h in stall_cols:
                    try:
                        st[h[6:]] += int(r[i])
                    except ValueError:
                        pass
            tot, ts = sum(ops.values()), sum(st.values())
            key_metrics[tag]["instruction_mix_pct"] = {k: round(100 * v / tot, 1) for k, v in ops.most_common(12)}
            key_metrics[tag]["stall_samples_pct"] = {k: round(100 * v / ts, 1) for k, v in st.most_common(8)}
            key_metrics[tag]["warp_instructions_per_variable"] = round(tot * 32 / 1e7, 1)
    if key_metrics:
        json.dump(key_metrics, open(os.path.join(P, f"{rnd}_ncu_key_metrics.json"), "w"), indent=1)
    if "ccsaq_m4" in summary:
        out = dict(summary["ccsaq_m4"])
        out["captures"] = summary
        out["note"] = "ncu --set full, one launch of dual_eval_kernel (CCSAQ, n=1e7, m=4, no x* store); per launch"
        json.dump(out, open(os.path.join(P, "ncu_summary.json"), "w"), indent=1)
    lb = os.path.join(G, "launches_bench.csv")
    if os.path.exists(lb):
        rows = [r for r in csv.reader(open(lb)) if len(r) > 5]
        hdr, data = None, []
        for r in rows:
            if r[0] == "ID":
                hdr = r
            elif hdr and r[0].isdigit():
                data.append(dict(zip(hdr, r)))
        agg = collections.defaultdict(lambda: [0, 0.0])
        for d in data:
            v = float(d["Metric Value"].replace(",", ""))
            v = v / 1e3 if d["Metric Unit"] == "ns" else (v * 1e3 if d["Metric Unit"] == "ms" else v)
            name = d["Kernel Name"].split("(")[0]
            agg[name][0] += 1
            agg[name][1] += v
        tot = sum(v[1] for v in agg.values())
        with open(os.path.join(P, f"{rnd}_launches_bench_share.txt"), "w") as f:
            f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu\n")
            f.write("# per-kernel totals over the whole process (cold-cache, serialised: compare SHARES)\n")
            for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write(f"{k[:90]:92s} launches={v[0]:5d} total_us={v[1]:11.1f} avg_us={v[1] / v[0]:9.1f} share={100 * v[1] / tot:5.1f}%\n")
        shutil.copy(lb, os.path.join(P, f"{rnd}_launches_bench.csv"))
    for f in ("kernel_sweep.json", "bench_n1.json"):
        if os.path.exists(os.path.join(G, f)):
            shutil.copy(os.path.join(G, f), os.path.join(P, f"{rnd}_{f}"))
    print(json.dumps(key_metrics, indent=1)[:3000])


if __name__ == "__main__":
    main()
