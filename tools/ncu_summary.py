"""Turn an .ncu-rep (one `--set full` capture of dla_loglik_kernel) into the small JSON/text summary that is
committed under profiles/ and read by bench.py for roofline.traffic."""
import csv, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_imma_cycles_active_realtime.avg", "sm__cycles_active.avg",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "launch__cluster_dim_x",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum"]
d = {}
for h, u, v in zip(hdr, units, vals):
    for w in want:
        if h == w or h.endswith("." + w):      # some metrics carry a section prefix (TPC.TriageCompute. ...)
            d[w] = {"value": v, "unit": u}
def num(k):
    return float(d[k]["value"].replace(",", ""))
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
traffic = sum(num(k) * scale[d[k]["unit"]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
d["dram_bytes_per_launch"] = traffic
json.dump(d, open(out, "w"), indent=1)
print(json.dumps({k: v for k, v in d.items() if k != "Kernel Name"}, indent=1))
