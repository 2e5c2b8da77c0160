"""Markdown summary of ncu reports: tools/ncu_summary.py title=report.ncu-rep ... > profiles/x.md"""
import csv, subprocess, sys
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "launch__cluster_size"]
for arg in sys.argv[1:]:
    title, rep = arg.split("=", 1)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"## {title}\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"kernel `{name[:90]}` (`{rep.split('/')[-1]}`)\n")
        print("| metric | value |\n|---|---|")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"| {w} | {r[i]} {units[i]} |")
        print()
