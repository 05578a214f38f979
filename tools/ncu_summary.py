"""ncu -i <rep> --page raw --csv  ->  a small CSV of the metrics the roofline discussion uses."""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_requests_srcunit_tex_op_read.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "dram__sectors_read.sum", "dram__sectors_write.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(WANT)
        w.writerow([units[hdr.index(k)] if k in hdr else "" for k in WANT])
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            w.writerow([d.get(k, "") for k in WANT])
    print(out, len(rows) - 2, "launches")


if __name__ == "__main__":
    main()
