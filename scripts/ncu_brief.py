"""Compact per-kernel summary of an .ncu-rep: python scripts/ncu_brief.py file.ncu-rep [units_per_launch]"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
units_row = dict(zip(hdr, rows[1]))  # ncu prints each raw metric in a unit of its own choosing
K = {
    "time_us": "gpu__time_duration.sum", "cycles": "sm__cycles_elapsed.max", "warp_inst": "smsp__inst_executed.sum",
    "thread_inst": "thread_inst_executed", "issue_active%": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "alu%": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "fma%": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "fmaheavy%": "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "lsu_wavefronts_shared": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "shared_ld": "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "shared_st": "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "dram_rd_MB": "dram__bytes_read.sum", "dram_wr_MB": "dram__bytes_write.sum", "dram%": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "regs": "launch__registers_per_thread", "occ_warps%": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "lim_smem": "launch__occupancy_limit_shared_mem", "lim_regs": "launch__occupancy_limit_registers",
}
ST = ["short_scoreboard", "long_scoreboard", "barrier", "math_pipe_throttle", "wait", "not_selected", "mio_throttle", "lg_throttle",
      "branch_resolving", "dispatch_stall", "no_instruction", "membar", "sleeping", "tex_throttle", "drain", "imc_miss", "selected"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("===", d["Kernel Name"][:70], "grid", d.get("launch__grid_size"), "block", d.get("launch__block_size"))
    line = []
    for k, m in K.items():
        if m in d and d[m] not in ("", "no data"):
            u = units_row.get(m, "") if k.startswith("dram_") or k == "time_us" else ""
            line.append(f"{k.replace('_MB', '')}={float(d[m].replace(',', '')):.4g}{('[' + u + ']') if u else ''}")
    print("  " + "  ".join(line))
    if units:
        cyc = float(d["sm__cycles_elapsed.max"].replace(",", ""))
        wi = float(d["smsp__inst_executed.sum"].replace(",", ""))
        sw = float(d["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"].replace(",", ""))
        print(f"  per unit: {cyc * 148 / units:.1f} SM-cycles, {wi / units:.1f} warp-instr, {sw / units:.1f} shared wavefronts")
    st = []
    for x in ST:
        m = f"smsp__average_warps_issue_stalled_{x}_per_issue_active.ratio"
        if m in d and d[m] not in ("", "no data") and float(d[m]) >= 0.05:
            st.append(f"{x}={float(d[m]):.2f}")
    print("  stalls/issue: " + " ".join(st))
