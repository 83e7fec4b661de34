"""Turn an `ncu --set full` capture into what the repository keeps: a text summary under profiles/ and the entry of
profiles/ncu_summary.json that bench.py quotes in `roofline` (DRAM bytes per read and the counters that name the
binding resource of the workload's dominant kernel).

usage: ncu_to_profiles.py report.ncu-rep|raw.csv workload reads_per_launch kernel_substring output_name
       (workload "-" : write the text summary only, e.g. for a kernel that is not its workload's dominant one)
"""
import csv, json, os, subprocess, sys

KEYS = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed")


def number(text):
    try:
        return float(text.replace(",", ""))
    except ValueError:
        return None


def main():
    rep, workload, reads, kernel, name = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4], sys.argv[5]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # either the .ncu-rep itself or the `ncu -i rep --page raw --csv` dump of it (what travels back from the GPU box)
    out = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    header, units = rows[0], rows[1]
    row = next(r for r in rows[2:] if kernel in r[header.index("Kernel Name")])
    value = {h: (row[i], units[i]) for i, h in enumerate(header)}
    lines = ["%s  (%s, %d reads per launch, ncu --set full --clock-control none)" % (row[header.index("Kernel Name")], os.path.basename(rep), reads)]
    for h in header:
        if h in KEYS or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")):
            lines.append("  %-90s %-16s %s" % (h, value[h][1], value[h][0]))
    text_path = os.path.join(root, "profiles", name + ".txt")
    open(text_path, "w").write("\n".join(lines) + "\n")

    def scaled(key):
        v, unit = value[key]
        v = number(v)
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        return v * scale
    dram = scaled("dram__bytes_read.sum") + scaled("dram__bytes_write.sum")
    percent = lambda key: number(value[key][0])
    candidates = {"issue slots (smsp__issue_active)": percent("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                  "ALU pipe (sm__inst_executed_pipe_alu)": percent("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                  "LSU pipe (sm__inst_executed_pipe_lsu)": percent("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
                  "shared-memory wavefronts (l1tex__data_pipe_lsu_wavefronts_mem_shared)": percent("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
                  "FP64 pipe (sm__inst_executed_pipe_fp64)": percent("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                  "DRAM (gpu__dram_throughput)": percent("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")}
    resource = max(candidates, key=lambda k: candidates[k] or 0)
    if workload == "-":
        print(open(text_path).read())
        return
    summary_path = os.path.join(root, "profiles", "ncu_summary.json")
    summary = json.load(open(summary_path)) if os.path.exists(summary_path) else {}
    summary[workload] = {"kernel": row[header.index("Kernel Name")], "dram_bytes_per_read": dram / reads,
                         "binding": {"resource": resource, "pct_of_peak": candidates[resource], "counters_pct_of_peak": candidates,
                                     "warp_instructions_per_32_reads": number(value["smsp__inst_executed.sum"][0]) / (reads / 32.0),
                                     "kernel_us": number(value["gpu__time_duration.sum"][0]) * {"us": 1, "ms": 1e3, "ns": 1e-3}.get(value["gpu__time_duration.sum"][1], 1)},
                         "source": "profiles/%s.txt (offline ncu capture, file constants)" % name}
    json.dump(summary, open(summary_path, "w"), indent=1)
    print(open(text_path).read())


if __name__ == "__main__":
    main()
