"""Turn an .ncu-rep into the small text artefacts committed under profiles/.
usage: python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r1_name
writes <out>_metrics.csv (selected raw metrics) and <out>_hotspots.txt (per-region stall mix)."""
import collections
import csv
import io
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor", "sm__pipe_tensor_cycles_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warps_issue_stalled", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep, out):
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    with open(out + "_metrics.csv", "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "metric", "unit", "value"])
        for vals in rows[2:]:
            name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            for h, u, v in zip(hdr, units, vals):
                if any(h.startswith(k) for k in KEEP) and ".max." not in h and ".min." not in h and ".sum.p" not in h:
                    w.writerow([name[:60], h, u, v])
    rows = ncu_csv(rep, "source")
    name, hdr, data = rows[0][1], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
    cls = collections.defaultdict(collections.Counter)
    ops = collections.Counter()
    for r in data:
        ie = int(r[ix["Instructions Executed"]] or 0)
        c = cls[ie]
        c["samples"] += int(r[ix["# Samples"]] or 0)
        c["ninstr"] += 1
        for s in stall_cols:
            c[s] += int(r[ix[s]] or 0)
        t = r[ix["Source"]].split()
        if t:
            op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
            ops[op.split(".")[0]] += ie
    with open(out + "_hotspots.txt", "w") as f:
        f.write(f"kernel: {name}\nwarp-stall samples: {tot}\n\n")
        f.write("code regions grouped by per-instruction execution count (= loop nest level):\n")
        for ie, c in sorted(cls.items(), key=lambda kv: -kv[1]["samples"])[:8]:
            top = sorted(((k, v) for k, v in c.items() if k.startswith("stall_")), key=lambda kv: -kv[1])[:6]
            f.write(f"  executed {ie:>10d} x {c['ninstr']:4d} instrs: {100 * c['samples'] / max(tot, 1):5.1f}% of samples; "
                    + ", ".join(f"{k[6:]} {100 * v / max(c['samples'], 1):.0f}%" for k, v in top) + "\n")
        total = sum(ops.values())
        f.write("\nwarp-level instruction mix:\n")
        for op, v in ops.most_common(14):
            f.write(f"  {op:10s} {v:>14d} {100 * v / total:5.1f}%\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
