"""summarise an ncu report (--set full) per kernel: duration, DRAM bytes, occupancy, top stall reasons.
usage: python tests/ncu_summary.py gpurun_out/x.ncu-rep   (runs `ncu -i ... --page raw --csv` here, no GPU needed)"""
import csv
import subprocess
import sys


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    u = dict(zip(hdr, units))
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        g = lambda k: d.get(k, "")
        print("=== %s  grid %s block %s" % (g("Kernel Name")[:70], g("launch__grid_size"), g("launch__block_size")))
        print("  time %s %s | dram rd %s %s wr %s %s | dram %% %s | sm thr %% %s" % (
            g("gpu__time_duration.sum"), u.get("gpu__time_duration.sum"), g("dram__bytes_read.sum"), u.get("dram__bytes_read.sum"),
            g("dram__bytes_write.sum"), u.get("dram__bytes_write.sum"), g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            g("sm__throughput.avg.pct_of_peak_sustained_elapsed")))
        print("  regs %s | occ limit regs %s smem %s | waves %s | warps active %% %s | issue active %% %s | eligible/cycle %s" % (
            g("launch__registers_per_thread"), g("launch__occupancy_limit_registers"), g("launch__occupancy_limit_shared_mem"),
            g("launch__waves_per_multiprocessor"), g("sm__warps_active.avg.pct_of_peak_sustained_active"),
            g("smsp__issue_active.avg.pct_of_peak_sustained_active"), g("smsp__warps_eligible.avg.per_cycle_active")))
        print("  inst %s | L1 st sectors %s ld sectors %s | smem bank conflicts %s | tensor pipe %% %s | L2 thr %% %s" % (
            g("smsp__inst_executed.sum"), g("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum"), g("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"),
            g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"), g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") or g("sm__inst_executed_pipe_tensor.sum"),
            g("lts__throughput.avg.pct_of_peak_sustained_elapsed")))
        st = {}
        for k, v in d.items():
            if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k:
                try:
                    st[k.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(v)
                except ValueError:
                    pass
        tot = sum(st.values()) or 1
        print("  stalls: " + ", ".join("%s %.0f%%" % (k, 100 * v / tot) for k, v in sorted(st.items(), key=lambda x: -x[1])[:7]))


if __name__ == "__main__":
    main(sys.argv[1])
