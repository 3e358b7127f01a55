"""Summarise `ncu --set full` captures (gpurun_out/*.ncu-rep) into a table of the metrics the roofline is judged on, and write
the per-launch DRAM traffic of the conv kernel to profiles/r02_traffic.json (bench.py puts it into roofline.traffic).
Usage: python tools/ncu_summary.py gpurun_out/r02_full_*.ncu-rep"""
import csv
import io
import json
import os
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "time_us",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "l1tex__m_xbar2l1tex_read_bytes.sum": "l2_to_sm_MB",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "sm__cycles_elapsed.max": "sm_cycles",
}
SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}


def read(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    rec = {"kernel": vals[hdr.index("Kernel Name")][:60] if "Kernel Name" in hdr else "?"}
    for i, h in enumerate(hdr):
        if h in WANT:
            v = float(vals[i].replace(",", "")) if vals[i] not in ("", "n/a") else float("nan")
            rec[WANT[h]] = v * SCALE.get(units[i], 1.0)
    return rec


def main():
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    traffic = {}
    print(f"{'capture':34s} {'time us':>8s} {'tensor%':>8s} {'sm%':>6s} {'dram rd MB':>10s} {'dram wr MB':>10s} {'dram%':>6s} {'L2->SM MB':>10s} {'grid':>5s}")
    for p in sys.argv[1:]:
        r = read(p)
        name = os.path.basename(p).replace("r02_full_", "").replace(".ncu-rep", "")
        print(f"{name:34s} {r.get('time_us', float('nan')):8.1f} {r.get('tensor_active_pct', float('nan')):8.1f} {r.get('sm_throughput_pct', float('nan')):6.1f} "
              f"{r.get('dram_read_MB', float('nan')):10.1f} {r.get('dram_write_MB', float('nan')):10.1f} {r.get('dram_pct', float('nan')):6.1f} "
              f"{r.get('l2_to_sm_MB', float('nan')):10.1f} {int(r.get('grid', 0)):5d}")
        traffic[name] = {"dram_bytes_read": int(r.get("dram_read_MB", 0) * 1e6), "dram_bytes_write": int(r.get("dram_write_MB", 0) * 1e6),
                         "time_us_under_ncu": round(r.get("time_us", 0), 1), "tensor_pipe_active_pct": round(r.get("tensor_active_pct", 0), 1)}
    with open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w") as fh:
        json.dump({"source": "ncu --set full --clock-control none, one launch per layer (tools/conv_probe.py shapes, bs 8 480x640), "
                             "dram__bytes_read.sum + dram__bytes_write.sum per launch", "per_launch": traffic}, fh, indent=1)


if __name__ == "__main__":
    main()
