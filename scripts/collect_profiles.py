#!/usr/bin/env python
"""Copy the evidence of one GPU profiling pass (scripts/gpu_profile.sh + gpu_quick.sh) from gpurun_out/ into profiles/:
launch lists + summaries, bench lines, key metrics and source hot spots of the full ncu captures, and
profiles/step_traffic.json (DRAM bytes per sampler step + kernel shares, read by bench.py).

    python scripts/collect_profiles.py <gpurun tag> <profiles label> [steps in the launch list = 5]
"""
import csv
import gzip
import io
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
KEYS = ["launch__grid_size", "launch__cluster_dim_z", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
STEP_KERNELS = ("conv_umma_kernel", "attn_umma_kernel", "tr_umma_kernel", "sampler_kernel", "pack_ncl_kernel", "attention_kernel")


def step_facts(path, steps):
    with gzip.open(path, "rt") as f:
        lines = [l for l in f if not l.startswith("==")]
    t, rd, wr, n = {}, 0.0, 0.0, 0
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for row in csv.DictReader(lines):
        k = re.sub(r"\(.*", "", row["Kernel Name"])
        short = next((s for s in STEP_KERNELS if s in k), None)
        if short is None:
            continue
        v, u, m = float(row["Metric Value"].replace(",", "")), row["Metric Unit"], row["Metric Name"]
        if m == "gpu__time_duration.sum":
            t[short] = t.get(short, 0.0) + (v / 1e3 if u.startswith("ns") else v)
            n += 1
        elif m == "dram__bytes_read.sum":
            rd += v * mult.get(u, 1)
        elif m == "dram__bytes_write.sum":
            wr += v * mult.get(u, 1)
    tot = sum(t.values())
    return {"dram_bytes_per_step": (rd + wr) / steps, "dram_read_bytes_per_step": rd / steps, "launches_per_step": n / steps,
            "conv_umma_share": "%.1f%%" % (100 * t.get("conv_umma_kernel", 0) / tot),
            "attn_umma_share": "%.1f%%" % (100 * t.get("attn_umma_kernel", 0) / tot),
            "serialised_us_per_step": tot / steps}


def main(tag, label, steps=5):
    os.makedirs(P, exist_ok=True)
    facts = {}
    for wl in ("config2", "config3"):
        src = os.path.join(G, "%s_launches_%s.csv.gz" % (tag, wl))
        if os.path.exists(src):
            shutil.copy(src, os.path.join(P, "%s_launches_%s.csv.gz" % (label, wl)))
            shutil.copy(src.replace(".csv.gz", "_summary.txt"), os.path.join(P, "%s_launches_%s_summary.txt" % (label, wl)))
            facts[wl] = step_facts(src, steps)
            facts[wl]["source"] = "profiles/%s_launches_%s.csv.gz (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum of `bench.py --steps 2 --warmup 3`; cold-cache, serialised)" % (label, wl)
        b = os.path.join(G, "%s_bench_%s.json" % (tag, wl))
        if os.path.exists(b) and os.path.getsize(b) > 0:
            shutil.copy(b, os.path.join(P, "%s_bench_%s.json" % (label, wl)))
    if facts:
        json.dump(facts, open(os.path.join(P, "step_traffic.json"), "w"), indent=1)
    for rep in sorted(os.listdir(G)):
        if not (rep.startswith(tag + "_full_") and rep.endswith(".ncu-rep")):
            continue
        name = rep[len(tag) + 1:-8]
        raw = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        out = ["# key metrics of %s (ncu --set full --clock-control none); one column per captured launch" % rep]
        if rows:
            h = rows[0]
            for k in ["Kernel Name"] + KEYS:
                if k in h:
                    i = h.index(k)
                    out.append("%-66s %-16s %s" % (k, rows[1][i], "  ".join(r[i][:40] for r in rows[2:])))
            st = {k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]: rows[2][h.index(k)]
                  for k in h if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")}
            top = sorted(st.items(), key=lambda kv: -float(kv[1] or 0))[:8]
            out.append("stall reasons (warps per issue-active cycle, first launch): " + ", ".join("%s %.2f" % (k, float(v)) for k, v in top))
        hot = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_lines.py"), os.path.join(G, rep), "0", "25"],
                             capture_output=True, text=True).stdout
        out.append("\n# SASS instructions ranked by warp-stall samples (samples, executions, instruction) -- first captured launch")
        out.append(hot)
        open(os.path.join(P, "%s_%s.txt" % (label, name)), "w").write("\n".join(out))
        # (the binary .ncu-rep stays in gpurun_out/: only the text digest is committed)
    for f in sorted(os.listdir(G)):  # every other bench line of the pass, the attention sweep, the codec launch list
        if f.startswith(tag + "_bench_") and f.endswith(".json") and os.path.getsize(os.path.join(G, f)) > 0:
            shutil.copy(os.path.join(G, f), os.path.join(P, label + f[len(tag):]))
        if f in (tag + "_attn_bench.jsonl", tag + "_launches_codec.csv.gz", tag + "_launches_codec_summary.txt",
                 tag + "_launches_codec_encode.csv.gz", tag + "_launches_codec_encode_summary.txt", tag + "_multigpu_check_n2.log"):
            shutil.copy(os.path.join(G, f), os.path.join(P, label + f[len(tag):]))
    for extra in ("tl_c2.txt", "tl_c3.txt"):
        src = os.path.join(G, "%s_%s" % (tag, extra))
        if os.path.exists(src):
            shutil.copy(src, os.path.join(P, "%s_timeline_%s" % (label, extra[3:])))
    print(json.dumps(facts, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 5)
