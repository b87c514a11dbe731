#!/usr/bin/env python3
"""Turn the ncu reports brought back in gpurun_out/ into the committed summaries under profiles/.

    python profiles/summarize.py r1            # reads gpurun_out/r1_*.ncu-rep, gpurun_out/launches_bench.csv

Writes profiles/<round>_<kernel>_raw.csv (selected raw metrics), profiles/<round>_<kernel>_top_stalls.txt
(source-page hot spots), profiles/<round>_launches.csv (per-kernel totals of the bench launch list) and
profiles/<round>_summary.md.
"""
import collections
import csv
import glob
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")

KEEP = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_elapsed.avg.per_second", "sm__cycles_active.avg",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__cluster_size", "smsp__inst_executed.sum",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def raw_rows(rep):
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    return rows[0], rows[1], rows[2:]


def top_stalls(rep, k=12):
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv"]))))
    if len(rows) < 3:
        return []
    hdr = rows[1]
    isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for r in rows[2:]:
        try:
            s = int(r[isamp])
        except (ValueError, IndexError):
            continue
        data.append((s, r))
    tot = sum(d[0] for d in data) or 1
    out = []
    for s, r in sorted(data, key=lambda x: -x[0])[:k]:
        st = sorted(((int(r[i] or 0), h) for i, h in stall), reverse=True)[:2]
        out.append(f"{100 * s / tot:5.1f}%  exec={r[iex]:>9}  {r[isrc].strip()[:72]:72s}  {st}")
    return out


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r1"
    md = [f"# ncu summaries, round {rnd[1:]}\n",
          "Captured on a B200 with `ncu --set full --clock-control none --import-source on` (one kernel instance, "
          "after the same command had exited 0 without ncu); durations are cold-cache and serialised, so they are "
          "NOT bench values -- bench.py times with CUDA events.  Regenerate with `python profiles/summarize.py`.\n"]
    for rep in sorted(glob.glob(os.path.join(SRC, f"{rnd}_*.ncu-rep"))):
        name = os.path.basename(rep)[len(rnd) + 1:-len(".ncu-rep")]
        hdr, units, rows = raw_rows(rep)
        if not rows:
            continue
        row = rows[-1]
        kn = row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else name
        sel = [(h, row[i], units[i]) for i, h in enumerate(hdr) if h in KEEP]
        with open(os.path.join(OUT, f"{rnd}_{name}_raw.csv"), "w") as f:
            w = csv.writer(f)
            w.writerow(["metric", "value", "unit"])
            w.writerows(sel)
        stalls = top_stalls(rep)
        with open(os.path.join(OUT, f"{rnd}_{name}_top_stalls.txt"), "w") as f:
            f.write(kn + "\n" + "\n".join(stalls) + "\n")
        d = {h: (v, u) for h, v, u in sel}
        g = lambda k: d.get(k, ("", ""))[0]
        md.append(f"\n## {name}: `{kn[:100]}`\n")
        md.append("| duration | SM clock | DRAM read / write | DRAM %peak | L2 %peak | tensor pipe %active | warps active % | issue active % | regs |")
        md.append("|---|---|---|---|---|---|---|---|---|")
        md.append(f"| {g('gpu__time_duration.sum')} {d.get('gpu__time_duration.sum', ('', ''))[1]} | {g('sm__cycles_elapsed.avg.per_second')} GHz | "
                  f"{g('dram__bytes_read.sum')} {d.get('dram__bytes_read.sum', ('', ''))[1]} / {g('dram__bytes_write.sum')} {d.get('dram__bytes_write.sum', ('', ''))[1]} | "
                  f"{g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')} | {g('lts__throughput.avg.pct_of_peak_sustained_elapsed')} | "
                  f"{g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')} | {g('sm__warps_active.avg.pct_of_peak_sustained_active')} | "
                  f"{g('smsp__issue_active.avg.pct_of_peak_sustained_active')} | {g('launch__registers_per_thread')} |")
        md.append("\nTop stall locations (source page):\n\n```")
        md += stalls[:8]
        md.append("```")
    lst = os.path.join(SRC, "launches_bench.csv")
    if os.path.exists(lst):
        rows = list(csv.reader(open(lst)))
        hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
        hdr = rows[hi]
        kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
        agg = collections.defaultdict(lambda: [0, 0.0])
        for r in rows[hi + 1:]:
            if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
                continue
            n = r[kn].split("(")[0]
            agg[n][0] += 1
            agg[n][1] += float(r[mv].replace(",", ""))
        tot = sum(v[1] for v in agg.values())
        with open(os.path.join(OUT, f"{rnd}_launches.csv"), "w") as f:
            w = csv.writer(f)
            w.writerow(["kernel", "launches", "total_us", "share_pct", "avg_us"])
            for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                w.writerow([k, v[0], round(v[1] / 1e3, 2), round(100 * v[1] / tot, 2), round(v[1] / v[0] / 1e3, 3)])
        md.append(f"\n## launch list of `bench.py --steps 2 --warmup 3 --no-cpu` ({rnd}_launches.csv)\n")
        md.append("Captured with `-k regex:fp8` (this library's kernels only; the benchmark's torch data-generation kernels "
                  "are filtered out).  Cold-cache and serialised: compare shares, not absolutes.\n")
        md.append("| kernel | launches | total µs | share | avg µs |\n|---|---|---|---|---|")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
            md.append(f"| `{k[:80]}` | {v[0]} | {v[1] / 1e3:.1f} | {100 * v[1] / tot:.1f}% | {v[1] / v[0] / 1e3:.2f} |")
    with open(os.path.join(OUT, f"{rnd}_summary.md"), "w") as f:
        f.write("\n".join(md) + "\n")
    print("wrote", os.path.join(OUT, f"{rnd}_summary.md"))


if __name__ == "__main__":
    main()
