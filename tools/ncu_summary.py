"""Per-launch table (time, DRAM bytes, throughput, occupancy, IPC, top stalls) from an `ncu --set full` report.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_kernels.md
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "sm__inst_executed.avg.per_cycle_elapsed", "lts__t_sector_hit_rate.pct",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
]
STALLS = "smsp__average_warps_issue_stalled_"   # prefix of the warp-state breakdown


def to_base(v, unit):
    """value -> base unit (bytes / seconds)"""
    mul = {"byte": 1, "Kbyte": 1e3, "byte/block": 1, "Kbyte/block": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1,
           "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1}
    return v * mul.get(unit, 1)


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    unit = dict(zip(hdr, units))
    print(f"# ncu --set full summary of `{rep.split('/')[-1]}` (per launch; cold cache, serialised)\n")
    print("| kernel | grid | block | regs | smem B | time us | DRAM rd MB | DRAM wr MB | DRAM GB/s | DRAM % peak | SM % | occ % | IPC | L2 hit % | top stalls (cycles/issue) |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for r in rows[2:]:
        d = dict(zip(hdr, r))

        def f(name):
            try:
                return to_base(float(d[name].replace(",", "")), unit[name])
            except Exception:
                return float("nan")
        t = f("gpu__time_duration.sum")
        rd, wr = f("dram__bytes_read.sum"), f("dram__bytes_write.sum")
        stalls = []
        for k in hdr:
            if k.startswith(STALLS) and k.endswith(".ratio") and "selected" not in k:
                try:
                    stalls.append((float(d[k]), k[len(STALLS):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        smem = f("launch__shared_mem_per_block_dynamic") + f("launch__shared_mem_per_block_static")
        name = d["Kernel Name"].replace("void ", "").split("(")[0]
        print(f"| `{name}` | {d['launch__grid_size']} | {d['launch__block_size']} | {d['launch__registers_per_thread']} "
              f"| {smem:.0f} | {t * 1e6:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {(rd + wr) / t / 1e9:.0f} "
              f"| {f('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} "
              f"| {f('sm__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} "
              f"| {f('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} "
              f"| {f('sm__inst_executed.avg.per_cycle_elapsed'):.2f} | {f('lts__t_sector_hit_rate.pct'):.0f} "
              f"| {', '.join(f'{n} {v:.1f}' for v, n in stalls[:3])} |")


if __name__ == "__main__":
    main()
