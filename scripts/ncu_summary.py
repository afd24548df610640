"""Summarise an .ncu-rep (one or more launches) into a small JSON under profiles/: the counters the north star names
(issue efficiency, divergence, L1/L2 hit rates, DRAM traffic) plus stall reasons and pipe utilisation.
Usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/<name>.json"""
import csv, json, subprocess, sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.avg.per_cycle_elapsed", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "smsp__sass_average_branch_targets_threads_uniform.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.sum", "lts__t_requests.sum", "lts__t_requests_srcunit_tex_op_read.sum", "lts__t_requests_srcunit_tex_op_write.sum", "lts__t_sector_op_read_hit_rate.pct",
    "lts__t_bytes.sum", "lts__t_bytes.sum.per_second", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_bytes.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sass__inst_executed_global_loads", "sass__inst_executed_shared_loads",
]
STALL = "smsp__average_warps_issue_stalled_"


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return s


rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
res = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    e = {"kernel": d.get("Kernel Name"), "grid": d.get("Grid Size"), "block": d.get("Block Size")}
    for k in KEYS:
        if k in d and d[k] != "":
            e[k] = num(d[k]); e[k + "__unit"] = units[hdr.index(k)]
    e["stalls_per_issue"] = {h[len(STALL):-len("_per_issue_active.ratio")]: num(d[h]) for h in hdr
                             if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and d[h] not in ("", "0")}
    if "lts__t_sectors.sum" in e:
        e["l2_bytes"] = 32.0 * e["lts__t_sectors.sum"]          # 32-byte sectors through the L2 tag stage
    res.append(e)
json.dump(res, open(out, "w"), indent=1)
print("wrote", out, len(res), "launches")
