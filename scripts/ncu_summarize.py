"""Turn an .ncu-rep into a small text summary (run on the GPU box: the reports are too large to bring back).

    python scripts/ncu_summarize.py gpurun_out/x.ncu-rep gpurun_out/x_summary.txt ["header line"]

Per profiled launch: the raw-page metrics the roofline discussion uses, then the warp-stall samples of the source page summed
per stall reason and per opcode class (where the issue slots went)."""
import collections, csv, io, re, subprocess, sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_dim_x", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg",
]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    lines = [sys.argv[3] if len(sys.argv) > 3 else "# " + rep, ""]
    raw = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr = next((r for r in raw if "Kernel Name" in r), None)
    if hdr:
        col = {n: i for i, n in enumerate(hdr)}
        units = raw[raw.index(hdr) + 1]
        for r in raw[raw.index(hdr) + 2:]:
            if len(r) < len(hdr):
                continue
            lines.append("-----\nKernel Name  " + r[col["Kernel Name"]])
            for m in WANT:
                if m in col:
                    lines.append("%-75s %s %s" % (m, r[col[m]], units[col[m]]))
            for n, i in col.items():
                if "issue_stalled" in n and n.endswith("per_issue_active.ratio"):
                    try:
                        if float(r[i]) >= 0.3:
                            lines.append("   %-85s %s" % (n, r[i]))
                    except ValueError:
                        pass
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv"]))))
    blocks, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": [], "hdr": None}; blocks.append(cur)
        elif cur is not None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] and len(r) >= len(cur["hdr"]):
            cur["rows"].append(r)
    done = set()
    for b in blocks:
        if b["hdr"] is None or not b["rows"]:
            continue
        key = (b["name"], len(b["rows"]), sum(int(r[b["hdr"].index("# Samples")] or 0) for r in b["rows"]))
        if key in done:                                     # ncu repeats the source page of a kernel for every launch of it
            continue
        done.add(key)
        h = b["hdr"]; col = {n: i for i, n in enumerate(h)}
        sc = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
        tot, byop, samples = collections.Counter(), collections.defaultdict(collections.Counter), 0
        for r in b["rows"]:
            s = re.sub(r"^@!?U?P\d+\s+", "", r[col["Source"]].strip())
            op = (s.split() or ["?"])[0].split(".")[0]
            samples += int(r[col["# Samples"]] or 0)
            for n in sc:
                v = int(r[col[n]] or 0)
                tot[n] += v; byop[op][n] += v
        lines.append("-----\nwarp-stall samples by reason: " + b["name"][:110] + "  (total %d)" % samples)
        for k, v in tot.most_common(12):
            lines.append("   %-28s %9d  %5.1f %%" % (k, v, 100.0 * v / max(1, samples)))
        lines.append("   top opcodes (samples: main reasons)")
        for op, c in sorted(byop.items(), key=lambda kv: -sum(kv[1].values()))[:10]:
            lines.append("   %-14s %9d  %s" % (op, sum(c.values()), ", ".join("%s %d" % (k.replace("stall_", ""), v) for k, v in c.most_common(3))))
    open(out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
