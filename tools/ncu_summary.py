"""Condense ncu exports into the small text files kept under profiles/.

    python tools/ncu_summary.py raw     <ncu --page raw --csv file>     <out.md>   [title]
    python tools/ncu_summary.py launches <ncu gpu__time_duration csv>    <out.csv>
    python tools/ncu_summary.py source  <ncu --page source --csv file>  <out.md>   [title]

`raw`      one block per profiled launch: duration, DRAM bytes (read+write = the bench's `roofline.traffic`), DRAM / SM
           throughput, occupancy, registers, issue utilisation and the warp-stall breakdown.
`launches` the launch list (id, kernel, grid, block, ns) with the template noise of torch's kernels shortened, plus
           a per-kernel share table appended as comment lines.
`source`   executed instructions and stall samples aggregated by SASS opcode (what the kernel spends its issue slots on).
"""
import collections
import csv
import re
import sys

csv.field_size_limit(10 ** 9)

RAW_KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu dram throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__shared_mem_per_block_static", "static smem / block"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers)"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem)"),
    ("launch__occupancy_limit_warps", "occupancy limit (warps)"),
    ("launch__waves_per_multiprocessor", "waves / SM"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]


def short_kernel(name):
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
    if len(name) > 110:
        name = name[:107] + "..."
    return name


def raw(path, out, title):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as fh:
        fh.write("# %s\n\nsource: `ncu --set full --clock-control none --import-source on`, read with "
                 "`ncu -i <rep> --page raw --csv` and condensed by `tools/ncu_summary.py raw`.\n"
                 "Durations under ncu are serialised, cold-cache replays: compare shares and ratios, not absolutes.\n\n"
                 % title)
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            fh.write("## %s  grid %s block %s\n\n" % (short_kernel(d.get("Kernel Name", "?")), d.get("Grid Size"),
                                                    d.get("Block Size")))
            fh.write("| metric | value |\n|---|---|\n")
            rd = wr = None
            for k, label in RAW_KEYS:
                if k in d and d[k] != "":
                    fh.write("| %s (`%s`) | %s %s |\n" % (label, k, d[k], u.get(k, "")))
                    if k == "dram__bytes_read.sum":
                        rd = (float(d[k]), u[k])
                    if k == "dram__bytes_write.sum":
                        wr = (float(d[k]), u[k])
            if rd and wr and rd[1] == wr[1]:
                fh.write("| **traffic = dram read + write** | %.3f %s |\n" % (rd[0] + wr[0], rd[1]))
            stalls = [(k.split("stalled_")[1].split("_per_")[0], float(d[k])) for k in hdr
                      if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and d.get(k)]
            if stalls:
                stalls.sort(key=lambda kv: -kv[1])
                fh.write("\nwarp stalls (warps per issue-active cycle): " +
                         ", ".join("%s %.2f" % kv for kv in stalls if kv[1] >= 0.05) + "\n")
            fh.write("\n")


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if r]
    start = next(i for i, r in enumerate(rows) if r[0] == "ID")
    hdr = rows[start]
    idx = {k: hdr.index(k) for k in ("ID", "Kernel Name", "Block Size", "Grid Size", "Metric Value")}
    per = collections.OrderedDict()
    total = 0.0
    with open(out, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["id", "kernel", "grid", "block", "gpu__time_duration.sum [ns]"])
        for r in rows[start + 1:]:
            if len(r) <= idx["Metric Value"]:
                continue
            name = short_kernel(r[idx["Kernel Name"]])
            ns = float(r[idx["Metric Value"]].replace(",", ""))
            w.writerow([r[idx["ID"]], name, r[idx["Grid Size"]], r[idx["Block Size"]], "%.0f" % ns])
            c = per.setdefault(name, [0, 0.0])
            c[0] += 1
            c[1] += ns
            total += ns
        fh.write("# per-kernel share of the summed device time (%d launches, %.1f us)\n" %
                 (sum(c[0] for c in per.values()), total / 1e3))
        for name, (n, ns) in sorted(per.items(), key=lambda kv: -kv[1][1]):
            fh.write("# %5.1f%%  %4d x  avg %8.2f us  %s\n" % (100 * ns / total, n, ns / n / 1e3, name))


def source(path, out, title):
    rows = list(csv.reader(open(path)))
    kern, hdr = None, None
    data = collections.OrderedDict()
    for r in rows:
        if r and r[0] == "Kernel Name":
            kern = r[1]
            data[kern] = []
        elif r and r[0] == "Address":
            hdr = r
        elif kern and len(r) > 5:
            data[kern].append(r)
    with open(out, "w") as fh:
        fh.write("# %s\n\nsource: `ncu -i <rep> --page source --csv` (SASS view), aggregated by opcode with "
                 "`tools/ncu_summary.py source`.\n\n" % title)
        for k, rs in data.items():
            i_inst, i_src, i_smp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
            tot, smp = collections.Counter(), collections.Counter()
            for r in rs:
                parts = r[i_src].split()
                op = (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0]
                tot[op] += int(r[i_inst] or 0)
                smp[op] += int(r[i_smp] or 0)
            T, S = sum(tot.values()) or 1, sum(smp.values()) or 1
            fh.write("## %s\n\n%d SASS instructions, %d warp instructions executed, %d stall samples\n\n"
                     "| opcode | executed | stall samples |\n|---|---|---|\n" % (short_kernel(k), len(rs), T, S))
            for op, c in tot.most_common(18):
                fh.write("| %s | %.1f%% | %.1f%% |\n" % (op, 100.0 * c / T, 100.0 * smp[op] / S))
            fh.write("\n")


if __name__ == "__main__":
    what, src, dst = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else src
    {"raw": lambda: raw(src, dst, title), "launches": lambda: launches(src, dst),
     "source": lambda: source(src, dst, title)}[what]()
