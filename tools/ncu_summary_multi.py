"""Per-launch table of `ncu --set full` captures with several launches each: time, DRAM bytes and achieved DRAM GB/s, DRAM / SM /
tensor-pipe utilisation, L2->SM bytes, grid, registers.   python tools/ncu_summary_multi.py a.ncu-rep b.ncu-rep ... > summary.txt"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__grid_size", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active"]
SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}


def main():
    print(f"{'kernel':44s} {'grid':>6s} {'regs':>4s} {'time us':>8s} {'rd MB':>8s} {'wr MB':>8s} {'DRAM GB/s':>9s} {'dram%':>6s} {'sm%':>6s} {'tensor%':>7s} {'L2->SM MB':>9s} {'warps%':>6s}")
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(f"{path}: empty")
            continue
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        for vals in rows[2:]:
            def get(h):
                i = col.get(h)
                if i is None or vals[i] in ("", "n/a"):
                    return float("nan")
                return float(vals[i].replace(",", "")) * SCALE.get(units[i], 1.0)
            name = vals[col["Kernel Name"]].replace("void ", "").replace("prn::", "")[:44]
            t, rd, wr = get(WANT[0]), get(WANT[1]), get(WANT[2])
            print(f"{name:44s} {int(get(WANT[7])):6d} {int(get(WANT[8])):4d} {t:8.1f} {rd:8.1f} {wr:8.1f} {(rd + wr) / t * 1e3 if t == t and t > 0 else float('nan'):9.0f} "
                  f"{get(WANT[3]):6.1f} {get(WANT[4]):6.1f} {get(WANT[5]):7.1f} {get(WANT[6]):9.1f} {get(WANT[9]):6.1f}")


if __name__ == "__main__":
    main()
