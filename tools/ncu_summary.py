"""Summarise an `ncu --set full` report into a small CSV (one row per metric) and a markdown digest for profiles/.

usage: python tools/ncu_summary.py <report.ncu-rep> <profiles/name> ["free-text note"]

Writes <profiles/name>.csv (the selected raw metrics, every profiled launch) and prints a markdown block
(appended by the caller to profiles/README.md).  Runs on the CPU box: `ncu -i` needs no GPU.
"""
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.sum", "sm__inst_executed.sum.per_cycle_elapsed", "smsp__inst_executed.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, launches = rows[0], rows[1], rows[2:]
    idx = {k: i for i, k in enumerate(hdr)}
    with open(out + ".csv", "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["metric", "unit"] + [f"launch{j}" for j in range(len(launches))])
        w.writerow(["Kernel Name", ""] + [r[idx["Kernel Name"]] for r in launches])
        for k in KEEP:
            if k in idx:
                w.writerow([k, units[idx[k]]] + [r[idx[k]] for r in launches])
    r = launches[0]

    def g(k, default=None):
        return r[idx[k]] if k in idx and r[idx[k]] != "" else default

    def f(k):
        v = g(k)
        return float(v.replace(",", "")) if v is not None else float("nan")

    def scale(k):
        u = units[idx[k]] if k in idx else ""
        return {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(u, 1.0)

    rd = f("dram__bytes_read.sum") * scale("dram__bytes_read.sum")
    wr = f("dram__bytes_write.sum") * scale("dram__bytes_write.sum")
    tu = units[idx["gpu__time_duration.sum"]]
    t_ms = f("gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6, "msecond": 1.0, "usecond": 1e-3, "second": 1e3, "nsecond": 1e-6}.get(tu, 1.0)
    print(f"### `{out.split('/')[-1]}` — {g('Kernel Name')}")
    if note:
        print(note)
    print(f"* source report: `{rep}` (scratch, not committed); selected raw metrics: `{out}.csv`")
    print(f"* grid {g('launch__grid_size')} x {g('launch__block_size')} threads, {g('launch__registers_per_thread')} regs/thread, "
          f"duration under ncu {t_ms:.3f} ms (cold cache, serialised; not a bench number)")
    print(f"* DRAM traffic per launch: read {rd/1e6:.2f} MB + write {wr/1e6:.2f} MB = **{(rd+wr)/1e6:.2f} MB** "
          f"({f('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.2f}% of ncu's DRAM peak)")
    print(f"* issue slots busy {f('sm__issue_active.avg.pct_of_peak_sustained_elapsed'):.1f}%, "
          f"FMA pipe {f('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed'):.1f}%, "
          f"FP64 pipe {f('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed'):.1f}%, "
          f"XU {f('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed'):.1f}%, "
          f"tensor pipe {f('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):.1f}% (of elapsed cycles); "
          f"warps active {f('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f}% of max")
    stalls = {k.split("stalled_")[1].split("_per_issue")[0]: f(k) for k in KEEP if "warps_issue_stalled" in k and k in idx}
    top = sorted(stalls.items(), key=lambda kv: -kv[1])[:5]
    print("* top stall reasons (warps stalled per issue-active cycle): " + ", ".join(f"{k} {v:.2f}" for k, v in top))
    print()


if __name__ == "__main__":
    main()
