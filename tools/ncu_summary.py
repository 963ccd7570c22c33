"""Text summary of an ncu report for profiles/: `python tools/ncu_summary.py report.ncu-rep [kernel-regex] > profiles/x.txt`.
Reads `ncu -i report --page raw --csv`; one column per captured launch whose name matches the regex."""
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum.per_second",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__inst_executed_op_tma_ld.sum",
    "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.sum",
    "sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.sum.per_second",
    "smsp__sass_inst_executed_op_tmem_ldt.sum",
]


def main():
    rep = sys.argv[1]
    rx = re.compile(sys.argv[2]) if len(sys.argv) > 2 else re.compile(".")
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name")
    data = [r for r in data if rx.search(r[kn])]
    if not data:
        sys.exit("no kernel matches")
    print("Kernel Name []: " + " | ".join(r[kn][:90] for r in data))

    def line(k):
        i = hdr.index(k)
        print(f"{k} [{units[i]}]: " + " | ".join(r[i] for r in data))
    for k in KEYS:
        if k in hdr:
            line(k)
    stalls = [k for k in hdr if "issue_stalled" in k and k.endswith("per_issue_active.ratio")]
    vals = sorted(((max(float(r[hdr.index(k)] or 0) for r in data), k) for k in stalls), reverse=True)
    for v, k in vals[:8]:
        line(k)
    if "dram__bytes_read.sum" in hdr:
        i, j = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
        tr = [float(r[i]) * mult[units[i]] + float(r[j]) * mult[units[j]] for r in data]
        print("dram traffic per launch (read+write) [byte]: " + str(tr))


if __name__ == "__main__":
    main()
