"""Print a compact table of the key ncu metrics of every kernel in a .ncu-rep (reads `ncu -i ... --page raw --csv`)."""
import csv, subprocess, sys

KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("lts__t_sector_hit_rate.pct", "l2_hit%"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("launch__registers_per_thread", "regs"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"), ("sm__cycles_elapsed.avg.per_second", "clk"),
        ("lts__t_sectors_op_read.sum", "l2_rd_sect"), ("lts__t_sectors_op_write.sum", "l2_wr_sect")]

def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        parts = [name[:60]]
        for k, short in KEYS:
            if k in hdr:
                i = hdr.index(k)
                parts.append(f"{short}={r[i]}{units[i]}")
        print("  ".join(parts))

if __name__ == "__main__":
    for p in sys.argv[1:]:
        print("==", p)
        main(p)
