"""Summarise ncu outputs for profiles/:
    python tools/ncu_summary.py launches <launches.csv>          -> per-kernel totals / shares (markdown)
    python tools/ncu_summary.py full <file.ncu-rep>              -> key metrics per captured launch (markdown)
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    start = rows.index(hdr)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        try:
            agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    tot = sum(sum(v) for v in agg.values())
    print("| kernel | launches | total ms | mean us | share of step |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"| `{k[:70]}` | {len(v)} | {sum(v) / 1e6:.3f} | {sum(v) / len(v) / 1e3:.1f} | {sum(v) / tot:.3f} |")
    print(f"\ntotal {tot / 1e6:.3f} ms over {sum(len(v) for v in agg.values())} launches (cold-cache, serialised: compare shares)")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    name_i = hdr.index("Kernel Name")
    for r in rows[2:]:
        print(f"### `{r[name_i]}`\n\n| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"| {k} | {r[i]} | {units[i]} |")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
