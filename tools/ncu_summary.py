"""Summarise ncu outputs into small text files for profiles/.
  python tools/ncu_summary.py launches <launches.csv>      -> per-kernel share table
  python tools/ncu_summary.py full <report.ncu-rep>        -> key metrics of each captured launch
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        # shared-memory wavefronts (128 B) fetched by the tensor core (MMA operands) and by load/store instructions
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(io.StringIO("".join(lines))):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        ms = v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else v)
        k = row["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        agg[k][0] += 1
        agg[k][1] += ms
    tot = sum(v[1] for v in agg.values())
    print("total device time of the captured launches: %.3f ms" % tot)
    print("%-44s %6s %10s %7s" % ("kernel", "n", "ms", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-44s %6d %10.3f %6.1f%%" % (k[:44], v[0], v[1], 100 * v[1] / tot))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("kernel:", r[name_col][:100])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("  %-95s %s %s" % (k, r[i], units[i]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
