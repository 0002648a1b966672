"""Summarise an ncu report (read here, no GPU): python tools/ncu_summary.py gpurun_out/x.ncu-rep [launch index] > profiles/x.txt"""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__cluster",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.sum ", "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tmem.avg.pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct", "smsp__inst_executed.sum ",
        "dram__bytes_read.sum ", "dram__bytes_write.sum ", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum ", "lts__t_sectors.sum ", "lts__t_sectors_srcunit_tex.sum ", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum ", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum ", "l1tex__m_l1tex2xbar_write_bytes_mem_global_op_tma_st.sum ",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum ", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum ", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum ",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ", "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum ", "l1tex__t_sector_hit_rate.pct", "sm__memory_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp", "smsp__warp_issue_stalled"]
rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else -1
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
head, units, data = rows[0], rows[1], rows[2:]
row = data[which]
print(f"# {rep}: {len(data)} captured launch(es); this is launch {which if which >= 0 else len(data) + which}")
for i, name in enumerate(head):
    if name in ("Kernel Name", "Block Size", "Grid Size"):
        print(f"{name}: {row[i]}")
for i, name in enumerate(head):
    if any((name + " ").startswith(k) or name.startswith(k) for k in KEYS) and row[i] not in ("", "0") and ".min" not in name and ".max." not in name and "per_second" not in name:
        print(f"{name} = {row[i]} {units[i]}")
