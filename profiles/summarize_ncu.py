"""Summarise an .ncu-rep (or a gpu__time_duration launch-list CSV) into the text files committed under profiles/.

  python profiles/summarize_ncu.py report gpurun_out/prof.ncu-rep      > profiles/rNN_kernels.txt
  python profiles/summarize_ncu.py launches gpurun_out/launches.csv    > profiles/rNN_launches.txt
"""
import collections
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]


def report(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H, units = rows[0], rows[1]
    stall = [h for h in H if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    print(f"# ncu --set full --clock-control none  ({path})")
    for r in rows[2:]:
        print("-----", r[H.index("Kernel Name")][:90])
        for w in WANT:
            if w in H:
                print(f"  {w:72s} {r[H.index(w)][:24]:>24s} {units[H.index(w)]}")
        try:
            rd = float(r[H.index("dram__bytes_read.sum")].replace(",", "")); wr = float(r[H.index("dram__bytes_write.sum")].replace(",", ""))
            print(f"  {'traffic = dram read + write (units as above)':72s} {rd:.6f} + {wr:.6f}")
        except ValueError:
            pass
        st = sorted(((float(r[H.index(h)].replace(",", "")), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                     for h in stall), reverse=True)
        print("  top stalls (warps per issue):", ", ".join(f"{n}={v:.2f}" for v, n in st[:8]))


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    agg = collections.OrderedDict()
    order = []
    for r in rows[hdr + 1:]:
        if len(r) > vi:
            agg.setdefault(r[ki][:70], []).append(float(r[vi].replace(",", "")))
            order.append((r[ki][:70], float(r[vi].replace(",", ""))))
    total = sum(sum(v) for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none ({path}); cold-cache serialised launches: compare SHARES")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:72s} n={len(v):4d} mean={sum(v)/len(v)/1e3:9.1f} us  share={100*sum(v)/total:5.1f}%")


if __name__ == "__main__":
    {"report": report, "launches": launches}[sys.argv[1]](sys.argv[2])
