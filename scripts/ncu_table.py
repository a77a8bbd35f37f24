"""Compact per-kernel table from an `ncu --page raw --csv` export (the handful of numbers DESIGN.md / profiles quote)."""
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]

    def g(r, k):
        return r[hdr.index(k)] if k in hdr else ""

    def f(r, k):
        v = g(r, k).replace(",", "")
        try:
            return f"{float(v):.3g}"
        except ValueError:
            return "-"
    print("id | kernel | time | grid | regs | dram read | dram write | dram% | L2% | xu% | fma% | issue% | tensor pipe active% | tensor operand (smem) pipe% | top stalls")
    for r in rows[2:]:
        name = g(r, "Kernel Name").replace("void unnamed>::", "").replace("unnamed>::", "")[:46]
        stalls = [(float(r[i] or 0), h) for i, h in enumerate(hdr)
                  if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
        st = ", ".join(f"{h.split('issue_stalled_')[1].split('_per')[0]}={v:.1f}" for v, h in sorted(stalls, reverse=True)[:3])
        u = lambda k: units[hdr.index(k)] if k in hdr else ""      # noqa: E731
        print(" | ".join([g(r, "ID"), name, f(r, "gpu__time_duration.sum") + " " + u("gpu__time_duration.sum"),
                          g(r, "launch__grid_size"), g(r, "launch__registers_per_thread"),
                          f(r, "dram__bytes_read.sum") + " " + u("dram__bytes_read.sum"),
                          f(r, "dram__bytes_write.sum") + " " + u("dram__bytes_write.sum"),
                          f(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                          f(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                          f(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                          f(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                          f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                          f(r, "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
                          f(r, "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                          st]))


if __name__ == "__main__":
    main(sys.argv[1])
