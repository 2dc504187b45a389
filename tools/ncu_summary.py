"""Summarise an `ncu --page raw --csv` export (one kernel launch) into the JSON kept under profiles/."""
import csv, json, sys
src, dst, label = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(src)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = {
    "gpu__time_duration.sum": "duration",
    "sm__cycles_elapsed.avg": "sm_cycles",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active": "tc_pipe_pct",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active": "tmem_pipe_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_lsu_wavefront_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__occupancy_limit_shared_mem": "occupancy_limit_smem_blocks",
    "launch__grid_size": "grid_size",
    "launch__block_size": "block_size",
    "smsp__sass_inst_executed_op_tmem_ldt.sum": "tmem_ld_instructions",
    "smsp__sass_inst_executed_op_tmem_stt.sum": "tmem_st_instructions",
}
out = {"kernel": label, "source": src.split("/")[-1], "how": "ncu --set full --clock-control none, one launch"}
for h, u, v in zip(hdr, units, vals):
    if h in want:
        try:
            out[want[h]] = {"value": float(v.replace(",", "")), "unit": u}
        except ValueError:
            out[want[h]] = {"value": v, "unit": u}
def scaled(key):
    d = out.get(key)
    if not d:
        return None
    mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(d["unit"], 1)
    return d["value"] * mul
r, w = scaled("dram_read"), scaled("dram_write")
if r is not None and w is not None:
    out["dram_bytes_per_launch"] = r + w
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out)[:400])
