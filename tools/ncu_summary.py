"""Markdown summary of an .ncu-rep (run where ncu is installed): one row per profiled launch with the metrics the design
discussion uses.  python tools/ncu_summary.py gpurun_out/x.ncu-rep "title" > profiles/x.md"""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time"), ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/smem data pipe %"),
        ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem: tensor-core operand reads %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem: LSU %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard / issue"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier / issue"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem/block"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %")]


def main():
    rep, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ik = hdr.index("Kernel Name")
    print("# %s\n" % title)
    print("`ncu --set full --clock-control none` (one launch per row; cold caches, serialised - shares and pipe utilisations are")
    print("the evidence, absolute times are not bench numbers).\n")
    cols = []
    for m, lab in WANT:
        hit = [i for i, h in enumerate(hdr) if h == m] or [i for i, h in enumerate(hdr) if m in h]
        if hit:
            cols.append((hit[0], lab))
    print("| kernel | " + " | ".join(lab for _, lab in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for r in data:
        name = r[ik].replace("mas::", "").replace("void ", "")
        name = name[:name.index("(")] if "(" in name else name
        cells = []
        for i, _ in cols:
            v = r[i]
            try:
                f = float(v.replace(",", ""))
                v = ("%.3g" % f) if abs(f) < 1000 else ("%.0f" % f)
            except ValueError:
                pass
            cells.append("%s %s" % (v, units[i] if units[i] not in ("%", "") else "%" if units[i] == "%" else ""))
        print("| `%s` | " % name + " | ".join(c.strip() for c in cells) + " |")


if __name__ == "__main__":
    main()
