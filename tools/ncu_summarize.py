"""Selected raw-page metrics of an .ncu-rep (ncu -i <rep> --page raw --csv) -> a small CSV for profiles/.
usage: python tools/ncu_summarize.py gpurun_out/prof_gemm.ncu-rep profiles/r01b_ncu_full_gemm.csv"""
import csv, io, subprocess, sys

KEEP = [
    "Kernel Name", "Block Size", "Grid Size",
    "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
]

def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    units = rows[1]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    # any tensor-pipe metric the report holds, whatever this ncu version calls it
    idx += [i for i, h in enumerate(hdr) if "pipe_tensor" in h and "pct_of_peak_sustained_active" in h and i not in idx]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            if len(r) == len(hdr):
                w.writerow([r[i] for i in idx])
    print(out, len(rows) - 2, "kernels,", len(idx), "metrics")

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
