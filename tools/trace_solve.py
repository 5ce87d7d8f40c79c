"""Timeline of the persistent dual-solve kernel: where does one generation (= one dual evaluation) spend its
time?  Builds an instrumented copy of the library (-DNB200_TRACE, build/trace/) and runs bench.py's
device-resident arm against it with NLOPT_B200_TRACE_FILE set.  Usage:
    python tools/trace_solve.py build            # here (no GPU needed)
    python tools/trace_solve.py run N [alg]      # on the GPU box; prints a summary of gpurun_out/trace_N.txt"""
import os
import re
import statistics
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as G  # noqa: E402

TDIR = os.path.join(ROOT, "build", "trace")


def build():
    os.makedirs(TDIR, exist_ok=True)
    cuda_home = os.path.dirname(os.path.dirname(G.NVCC))
    objs = []
    for src in G.LIB_SOURCES_CU:
        obj = os.path.join(TDIR, src + ".o")
        G._run([G.NVCC, *G.ARCH, *G.NVCC_FLAGS, "-DNB200_TRACE", "-c", os.path.join(G.CSRC, src), "-o", obj])
        objs.append(obj)
    for src in G.LIB_SOURCES_CXX:
        obj = os.path.join(TDIR, src + ".o")
        G._run(["g++", *G.CXX_FLAGS, f"-I{cuda_home}/include", "-c", os.path.join(G.CSRC, src), "-o", obj])
        objs.append(obj)
    lib = os.path.join(TDIR, "libnlopt_b200.so")
    G._run([G.NVCC, *G.ARCH, "-shared", "-o", lib, *objs, "-cudart", "shared", "-ldl", "-Xlinker", "-soname,libnlopt_b200.so",
            "-Xlinker", "-Bsymbolic-functions"])
    G._run([G.NVCC, *G.ARCH, *G.NVCC_FLAGS, "-shared", os.path.join(G.CSRC, "problems.cu"), "-o",
            os.path.join(TDIR, "libnlopt_b200_problems.so"), "-cudart", "shared", "-L" + TDIR, "-lnlopt_b200", "-Xlinker", "-rpath=$ORIGIN"])
    print("built", lib)


def run(n, alg="ccsaq", extra=()):
    out = os.path.join(ROOT, "gpurun_out", f"trace_{alg}_{n}.txt")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if os.path.exists(out):
        os.remove(out)
    env = dict(os.environ, NLOPT_B200_LIBDIR=TDIR, NLOPT_B200_TRACE_FILE=out)
    subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--n", str(n), "--alg", alg, "--steps", "4", "--warmup", "2",
                    "--no-cpu", "--no-e2e", *extra], env=env, check=True, stdout=subprocess.DEVNULL)
    cols = {k: [] for k in ("seen_lo", "seen_hi", "rec_lo", "rec_hi", "rank_done", "totals", "machine", "next_pub", "sweep")}
    head = ""
    pat = re.compile(r"gen \d+ .*?seen\[(-?\d+)\.\.(-?\d+)\] recs\[(-?\d+)\.\.(-?\d+)\] rank_done (-?\d+) totals (-?\d+) machine (-?\d+) next_pub (-?\d+) \| mean_group_sweep (\d+)")
    for line in open(out):
        if line.startswith("solve"):
            head = line.strip()
        mt = pat.search(line)
        if mt and int(mt.group(8)) > 0:
            for k, v in zip(cols, mt.groups()):
                cols[k].append(int(v))
    print(head)
    print(f"n={n} {alg}: {len(cols['sweep'])} generations; medians in ns after the generation was published:")
    for k, v in cols.items():
        print(f"  {k:10s} {statistics.median(v) if v else -1:9.0f}")


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    else:
        run(int(float(sys.argv[2])), *(sys.argv[3:4] or ["ccsaq"]), extra=sys.argv[4:])
