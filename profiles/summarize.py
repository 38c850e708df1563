"""Turn an ncu report (gpurun_out/*.ncu-rep) into the small CSV summary committed under profiles/.

    python profiles/summarize.py gpurun_out/prof.ncu-rep > profiles/r1_<name>.csv
"""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
        "smsp__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor"]
STALL = "smsp__average_warps_issue_stalled_"


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    w = csv.writer(sys.stdout)
    w.writerow(["launch", "metric", "value", "unit"])
    for k, r in enumerate(rows[2:]):
        for name in WANT:
            if name in hdr:
                i = hdr.index(name)
                w.writerow([k, name, r[i], units[i]])
        # the five largest warp-stall reasons (warps stalled per issued instruction)
        stalls = sorted(((float(r[i]), n) for i, n in enumerate(hdr)
                         if n.startswith(STALL) and n.endswith("_per_issue_active.ratio") and r[i]), reverse=True)
        for v, n in stalls[:5]:
            w.writerow([k, n, v, "warps/issue"])


if __name__ == "__main__":
    main(sys.argv[1])
