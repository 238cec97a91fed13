"""Text summary of an .ncu-rep (read here with `ncu -i`, no GPU needed): per profiled launch the metrics the roofline
discussion uses - duration, tensor-pipe / FP64 / issue activity, DRAM bytes and throughput, L2 and shared-memory
throughput, occupancy, launch geometry - plus the top stall reasons.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/r2_ncu_xxx.txt
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep --dram-csv profiles/r2_ncu_conv_dram.csv   # what bench.py reads
"""
import argparse
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"), ("launch__occupancy_limit_shared_mem", "occ limit smem"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % (elapsed)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % of active"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe % of elapsed"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor (hmma subpipe) % of active"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe % of active"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe cycles % of active"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipe % of active"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu pipe % of active"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu pipe % of active"),
    ("sm__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active % (smsp)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % (dram)"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/TEX throughput %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared memory wavefronts %"),
    ("smsp__cycles_active.avg", "active cycles per smsp"),
    ("sm__cycles_elapsed.max", "elapsed cycles"),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--dram-csv", default=None)
    ap.add_argument("--kernel", default=None, help="only launches whose name contains this")
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    rows = [r for r in rows if r and not r[0].startswith("==")]
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(hdr)}
    name_i = col.get("Kernel Name")
    if args.dram_csv:
        keep = ["ID", "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"]
        with open(args.dram_csv, "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(keep)
            w.writerow([units[col[k]] for k in keep])
            for r in data:
                if args.kernel and args.kernel not in r[name_i]:
                    continue
                w.writerow([r[col[k]] for k in keep])
    stall = [(n, i) for n, i in col.items() if n.startswith("smsp__average_warp") and "issue_stalled" in n and n.endswith("_per_issue_active.ratio")]
    if not stall:
        stall = [(n, i) for n, i in col.items() if "warp_issue_stalled" in n and n.endswith(".ratio")]
    for r in data:
        if args.kernel and args.kernel not in r[name_i]:
            continue
        print("== %s" % r[name_i][:150])
        for metric, label in WANT:
            if metric in col and r[col[metric]] != "":
                print("  %-36s %14s %s" % (label, r[col[metric]], units[col[metric]]))
        tops = []
        for n, i in stall:
            try:
                tops.append((float(r[i].replace(",", "")), n))
            except ValueError:
                pass
        tops.sort(reverse=True)
        for v, n in tops[:6]:
            short = n.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", "").replace("_per_issue_active.ratio", "")
            print("  stall %-30s %14.2f warps per issue-active cycle" % (short, v))
        print()


if __name__ == "__main__":
    main()
