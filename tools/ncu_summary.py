"""Summarise an .ncu-rep (ncu --set full) into a small text table for profiles/ (run where ncu is installed).
usage: python tools/ncu_summary.py report.ncu-rep > profiles/summary.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM bytes"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % active"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep}: one block per profiled launch (ncu --set full --clock-control none)")
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        short = name[name.find("::") + 2:] if "::" in name else name
        print(f"\n{short[:110]}  grid={r[idx['Grid Size']]} block={r[idx['Block Size']]}")
        for key, label in WANT:
            if key in idx:
                print(f"  {label:24s} {r[idx[key]]:>14s} {units[idx[key]]}")


if __name__ == "__main__":
    main()
