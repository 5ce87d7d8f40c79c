"""Turn the round-2 exports under gpurun_out/ into tracked evidence under profiles/ (run after tools/final_validation.sh and
tools/_gpu_mg.sh):  ncu key metrics of the persistent solve kernel (CCSAQ and MMA, n = 1e7, m = 4, 21 generations per
launch), profiles/ncu_summary.json (what bench.py reads), the launch-share table, copies of the bench / sweep lines."""
import collections
import csv
import glob
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
RND = sys.argv[1] if len(sys.argv) > 1 else "r02"      # evidence prefix: r02 (first session), r02b (second session)
KEYS = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio")


def num(v):
    return float(v.replace(",", ""))


def to_bytes(v, unit):
    return num(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    os.makedirs(P, exist_ok=True)
    gens = 21
    key = {}
    for tag, f in (("ccsaq_m4", "prof_solve_raw.csv"), ("mma_m4", "prof_solve_mma_raw.csv")):
        raw = os.path.join(G, f)
        if not os.path.exists(raw) or os.path.getsize(raw) < 100:
            continue
        rows = list(csv.reader(open(raw)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        k = {"kernel": d.get("Kernel Name", ("", ""))[0][:120], "generations_per_launch": gens}
        for kk in KEYS:
            if kk in d:
                k[kk] = f"{d[kk][0]} {d[kk][1]}".strip()
        rd, wr = to_bytes(*d["dram__bytes_read.sum"]), to_bytes(*d["dram__bytes_write.sum"])
        dur = num(d["gpu__time_duration.sum"][0]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(d["gpu__time_duration.sum"][1], 1.0)   # -> us
        k["dram_bytes_per_evaluation"] = (rd + wr) / gens
        k["algorithmic_bytes_per_evaluation"] = 8 * 10_000_000 * 9
        k["duration_us_per_evaluation_under_ncu"] = dur / gens
        if "smsp__inst_executed.sum" in d:
            k["warp_instructions_per_variable"] = num(d["smsp__inst_executed.sum"][0]) * 32 / (1e7 * gens)
        key[tag] = k
    if key:
        json.dump(key, open(os.path.join(P, f"{RND}_ncu_solve_key_metrics.json"), "w"), indent=1)
    if "ccsaq_m4" in key:
        c = key["ccsaq_m4"]
        out = {"dram_bytes_per_launch": c["dram_bytes_per_evaluation"], "algorithmic_bytes": c["algorithmic_bytes_per_evaluation"],
               "duration_us_under_ncu": c["duration_us_per_evaluation_under_ncu"],
               "note": "per dual evaluation (generation) of dual_solve_kernel, CCSAQ n=1e7 m=4: ncu --set full on one launch of 21 generations "
                       "(bench.py --param dual_maxeval=20), DRAM read+write bytes / 21; profiles/" + RND + "_ncu_solve_key_metrics.json", "captures": key}
        json.dump(out, open(os.path.join(P, "ncu_summary.json"), "w"), indent=1)
        for f, o in (("prof_solve_details.csv", f"{RND}_ncu_solve_kernel_details.csv"),):
            if os.path.exists(os.path.join(G, f)):
                shutil.copy(os.path.join(G, f), os.path.join(P, o))
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
            v = num(d["Metric Value"])
            v = v / 1e3 if d["Metric Unit"] == "ns" else (v * 1e3 if d["Metric Unit"] == "ms" else v)
            agg[d["Kernel Name"].split("(")[0]][0] += 1
            agg[d["Kernel Name"].split("(")[0]][1] += v
        tot = sum(v[1] for v in agg.values())
        with open(os.path.join(P, f"{RND}_launches_bench_share.txt"), "w") as fh:
            fh.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400  python bench.py --steps 2 --warmup 1 --no-cpu --no-parity\n")
            fh.write("# per-kernel totals over the whole process (device arm + e2e arm; cold-cache, serialised: compare SHARES)\n")
            for kname, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                fh.write(f"{kname[:90]:92s} launches={v[0]:5d} total_us={v[1]:11.1f} avg_us={v[1] / v[0]:9.1f} share={100 * v[1] / tot:5.1f}%\n")
        shutil.copy(lb, os.path.join(P, f"{RND}_launches_bench.csv"))
    for f in glob.glob(os.path.join(G, "bench_*.json")) + glob.glob(os.path.join(G, "sweep_c5_n*.json")):
        if os.path.getsize(f) > 10:
            shutil.copy(f, os.path.join(P, f"{RND}_{os.path.basename(f)}"))
    for f in (f"{RND}_final_n1.log",):
        if os.path.exists(os.path.join(G, f)):
            shutil.copy(os.path.join(G, f), os.path.join(P, f.replace(".log", ".txt")))
    print(json.dumps(key, indent=1)[:2500])


if __name__ == "__main__":
    main()
