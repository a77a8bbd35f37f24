"""Summarise an .ncu-rep (ncu --set full) into the handful of numbers DESIGN.md / profiles/ quote."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        print(f"# {path}")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            print(f"kernel: {name[:110]}")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    print(f"  {k:100s} {r[i]:>16s} {units[i]}")
            stalls = [(float(r[i] or 0), h) for i, h in enumerate(hdr)
                      if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
            for v, h in sorted(stalls, reverse=True)[:5]:
                print(f"  stall {h.split('issue_stalled_')[1].split('_per')[0]:40s} {v:10.2f} warps/issue")


if __name__ == "__main__":
    main()
