#!/usr/bin/env python
"""Turn the raw files one `profiles/gpu_session.sh <tag>` call left in gpurun_out/ into the
tracked summaries under profiles/:

    profiles/<tag>_launches.csv     the ncu launch list (gpu__time_duration per launch)
    profiles/<tag>_bench.json       the bench line of the same session (not under a profiler)
    profiles/<tag>_summary.md       per-kernel shares of the step + the ncu --set full metrics
                                    the roofline numbers are read from

usage: python profiles/summarize.py <tag>      (needs `ncu` on PATH; no GPU)
"""
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "DRAM read % of peak"),
    ("dram__bytes_write.sum.pct_of_peak_sustained_elapsed", "DRAM write % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per warp instruction"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue slots busy %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe (POPC) %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("smsp__sass_inst_executed_op_local_ld.sum", "local loads"),
    ("smsp__sass_inst_executed_op_local_st.sum", "local stores"),
    ("smsp__average_warp_latency_per_inst_issued.ratio", "warp latency per instruction (cycles)"),
]


def raw_page(rep):
    if rep.endswith(".csv"):
        txt = open(rep).read()
    else:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    return [OrderedDict((h, (v, u)) for h, v, u in zip(hdr, r, units)) for r in rows[2:]]


def launches(path):
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", ""))))
    return rows


def short(name):
    return re.sub(r"\(.*", "", name).replace("void ", "")


def main():
    tag = sys.argv[1]
    md = [f"# ncu summary `{tag}`", "",
          f"Produced by `profiles/gpu_session.sh {tag}` (round 2: `profiles/gpu_session_r2.sh {tag}`, `profiles/gpu_session_r2t.sh {tag}`) on one B200 (through gpurun) and "
          f"`profiles/summarize.py {tag}` here.  Launch times under ncu are cold-cache and serialised: "
          "compare SHARES with the CUDA-event numbers of the bench line, not absolutes.", ""]
    bj = os.path.join(OUT, f"{tag}_bench.json")
    if os.path.exists(bj) and os.path.getsize(bj):
        shutil.copy(bj, os.path.join(PROF, f"{tag}_bench.json"))
        b = json.loads(open(bj).read().strip().splitlines()[-1])
        md += ["## bench line of the same session (CUDA events, no profiler)", "",
               f"* value {b['value']:.0f} {b['unit']} (device-resident), e2e {b['e2e']['value']:.0f} {b['unit']}, "
               f"{b['ms_per_step']:.2f} ms/step, clocks {b.get('clocks')}",
               f"* roofline: {json.dumps(b.get('roofline'))}",
               f"* dp: {json.dumps(b.get('dp'))}",
               f"* cpu_baseline: {json.dumps(b.get('cpu_baseline'))}", ""]
    rj = os.path.join(OUT, f"{tag}_bench_reference.json")
    if os.path.exists(rj) and os.path.getsize(rj):
        shutil.copy(rj, os.path.join(PROF, f"{tag}_bench_reference.json"))
    lc = os.path.join(OUT, f"{tag}_launches.csv")
    if os.path.exists(lc):
        shutil.copy(lc, os.path.join(PROF, f"{tag}_launches.csv"))
        rows = launches(lc)
        agg = OrderedDict()
        for k, ns in rows:
            if "at_cuda_detail" in k:                       # torch's own kernels (index build, batch simulation): not part of a step
                continue
            a = agg.setdefault(short(k), [0, 0.0])
            a[0] += 1
            a[1] += ns
        upload = ("relayout", "seed_build", "s3_isa_kernel", "search_kernel<1", "random_sector")      # index upload, the rank-count pass, the probe
        tot = sum(v[1] for k, v in agg.items() if not any(u in k for u in upload))
        md += ["## launch list (`ncu --metrics gpu__time_duration.sum`, bench.py --steps 2 --warmup 1)", "",
               "| kernel | launches | avg ms | share of the steps' kernel time (index upload and the rank-count pass excluded) |", "|---|---|---|---|"]
        for k, (n, ns) in agg.items():
            sh = "-" if any(u in k for u in upload) else f"{100 * ns / tot:.1f} %"
            md.append(f"| `{k}` | {n} | {ns / n / 1e6:.3f} | {sh} |")
        md.append("")
    def gbytes(d):
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            v, u = d.get(k, ("0", "byte"))
            tot += float(v.replace(",", "")) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        return tot
    traffic = {"capture": tag, "unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)"}
    # round 2: one capture of a whole chain step, and the search launch of config 2
    for what, key in (("chain", None), ("search_k4", "search_launch_k4")):
        rep = os.path.join(OUT, f"{tag}_{what}.ncu-rep")
        if not os.path.exists(rep):
            rep = os.path.join(OUT, f"{tag}_{what}_raw.csv")          # the session leaves the raw page as CSV (the reports are too big to bring back)
        if not os.path.exists(rep):
            continue
        md += [f"## `ncu --set full` : {what} (one step; launches of one kernel summed)", ""]
        pages = raw_page(rep)
        per = OrderedDict()
        for d in pages:
            nm = short(d['Kernel Name'][0])
            nm = nm.split("<")[0] + ("<" + nm.split("<")[1] if "search_kernel<" in nm else "")
            per.setdefault(nm, []).append(d)
        if key:
            traffic[key] = sum(gbytes(d) for d in pages)
        else:
            for nm, ds in per.items():
                traffic[nm] = sum(gbytes(d) for d in ds)
            traffic["search_launch"] = sum(gbytes(d) for d in pages if any(x in d['Kernel Name'][0] for x in ("s3_search", "s3_heavy", "s3_isbad")))
            traffic["chain_step"] = sum(gbytes(d) for d in pages)
        md += ["| kernel | launches | ms (under ncu) | DRAM read + write | L2 hit rate | registers | warp instructions | issue slots busy % | ALU pipe % | top stalls |",
               "|---|---|---|---|---|---|---|---|---|---|"]
        for nm, ds in per.items():
            d = max(ds, key=lambda x: float(x["gpu__time_duration.sum"][0].replace(",", "")))
            scale = {"nsecond": 1e-6, "ns": 1e-6, "usecond": 1e-3, "us": 1e-3, "msecond": 1.0, "ms": 1.0, "second": 1e3, "s": 1e3}
            ms = sum(float(x["gpu__time_duration.sum"][0].replace(",", "")) * scale.get(x["gpu__time_duration.sum"][1], 1e-6) for x in ds)
            stalls = [(k.split("stalled_")[1].replace("_per_issue_active.ratio", ""), float(v[0] or 0))
                      for k, v in d.items() if "issue_stalled" in k and k.endswith("per_issue_active.ratio")]
            stalls.sort(key=lambda x: -x[1])
            g = lambda k: d[k][0] if k in d else "-"
            md.append(f"| `{nm}` | {len(ds)} | {ms:.3f} | {sum(gbytes(x) for x in ds) / 1e9:.3f} GB | {g('lts__t_sector_hit_rate.pct')} | {g('launch__registers_per_thread')} | "
                      f"{g('smsp__inst_executed.sum')} | {g('sm__issue_active.avg.pct_of_peak_sustained_elapsed')} | {g('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active')} | "
                      + ", ".join(f"{k} {v:.2f}" for k, v in stalls[:4]) + " |")
        md.append("")
    for what in ("search", "dp"):
        rep = os.path.join(OUT, f"{tag}_{what}.ncu-rep")
        if not os.path.exists(rep):
            continue
        md += [f"## `ncu --set full` : {what}", ""]
        pages = raw_page(rep)
        for d in pages:
            traffic[short(d['Kernel Name'][0]).split("<")[0] + ("<" + short(d['Kernel Name'][0]).split("<")[1] if "search_kernel<" in d['Kernel Name'][0] else "")] = gbytes(d)
        if what == "search":
            # one search launch = the straight-line kernel + the enumerating kernel and its split path
            traffic["search_launch"] = sum(gbytes(d) for d in pages)
        for d in pages:
            md += [f"### `{short(d['Kernel Name'][0])}`", "", "| metric | value |", "|---|---|"]
            for k, label in KEYS:
                if k in d:
                    md.append(f"| {label} (`{k}`) | {d[k][0]} {d[k][1]} |")
            stalls = [(k.split("stalled_")[1].replace("_per_issue_active.ratio", ""), float(v[0] or 0))
                      for k, v in d.items() if "issue_stalled" in k and k.endswith("per_issue_active.ratio")]
            stalls.sort(key=lambda x: -x[1])
            md.append("| top stall reasons (warps per issue) | " + ", ".join(f"{k} {v:.2f}" for k, v in stalls[:5]) + " |")
            md.append("")
    if os.path.exists(bj) and os.path.getsize(bj):
        b = json.loads(open(bj).read().strip().splitlines()[-1])
        if b.get("l2_persistence_experiment"):
            md += ["## L2 access-policy windows over the index arrays (`S3_L2_EXPERIMENT=1`, CUDA events, 5 search launches each)", "",
                   "| window (largest the device allows, persisting carve-out at its maximum) | straight-line kernel ms | enumerator ms | search launch ms |", "|---|---|---|---|"]
            for k, v in b["l2_persistence_experiment"].items():
                md.append(f"| {k} | {v['easy_kernel_ms']:.3f} | {v['enumerator_ms']:.3f} | {v['search_launch_ms']:.3f} |")
            md += ["", "No window changes the launch by more than its run-to-run spread: reads are uniform over the genome, the seed tables replace the "
                   "top of the BWT, and nothing in the 56 GB index is touched often enough to be worth a carve-out of the 126 MB L2.", ""]
    # a capture that also holds the deep-DP stage sums launches of different sizes per kernel: not a per-launch figure, bench.py keeps the chain-only one
    if len(traffic) > 2 and not any("s3_stage" in k or "s3_csr" in k for k in traffic):
        # measured once on the probe kernel itself (profiles/r2k_probe_ncu.csv): DRAM bytes per independent random 32-byte load; kept across captures
        old_path = os.path.join(PROF, "ncu_traffic.json")
        keep = json.load(open(old_path)).get("random_sector_probe_dram_bytes_per_load") if os.path.exists(old_path) else None
        traffic["random_sector_probe_dram_bytes_per_load"] = keep or 120.6
        k4 = json.load(open(old_path)).get("search_launch_k4") if os.path.exists(old_path) else None
        if k4 and "search_launch_k4" not in traffic:
            traffic["search_launch_k4"] = k4                        # (from the capture of the k <= 4 launch, kept across chain captures)
        json.dump(traffic, open(os.path.join(PROF, "ncu_traffic.json"), "w"), indent=1)      # bench.py's roofline.traffic
    open(os.path.join(PROF, f"{tag}_summary.md"), "w").write("\n".join(md) + "\n")
    print("\n".join(md))


if __name__ == "__main__":
    main()
