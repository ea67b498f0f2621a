"""ncu -i <rep> --page raw --csv  ->  compact JSON of the counters the roofline discussion uses, one entry per captured launch.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/ncu_r02_x.json "how it was captured" """
import csv, io, json, subprocess, sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1tex__t_bytes.sum",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "sm__cycles_elapsed.max", "sm__cycles_active.avg", "smsp__cycles_active.avg", "gpc__cycles_elapsed.max",
    "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__sass_inst_executed_op_shared_ld.sum",
]


def main():
    rep, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kernels = []
    for r in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, r):
            if h == "Kernel Name":
                d[h] = v
            elif h in KEEP:
                d[f"{h} [{u}]" if u else h] = v
        kernels.append(d)
    json.dump({"source": note, "kernels": kernels}, open(out, "w"), indent=1)
    for k in kernels:
        print(k.get("Kernel Name", "?")[:40], {a.split(" ")[0].split(".")[0]: b for a, b in list(k.items())[1:6]})


if __name__ == "__main__":
    main()
