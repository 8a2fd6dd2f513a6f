"""Summarise an ncu report (raw page CSV) into the few numbers the roofline discussion needs.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/xxx.md"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]
print(f"# ncu summary of `{rep}`\n")
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print(f"## {name[:110]}\n")
    print("| metric | value | unit |\n|---|---|---|")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"| {k} | {r[i]} | {units[i]} |")
    st = []
    for i, hn in enumerate(hdr):
        if "issue_stalled" in hn and hn.endswith("per_issue_active.ratio") and r[i] not in ("", "0"):
            try:
                st.append((float(r[i]), hn.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    st.sort(reverse=True)
    print("\nwarp stalls per issued instruction: " + ", ".join(f"{n} {v:.2f}" for v, n in st[:8]) + "\n")

# optional second argument: also write the DRAM traffic of the FIRST kernel of the report as JSON (what bench.py reads
# for `roofline.traffic`): python tools/ncu_summary.py rep.ncu-rep profiles/ncu_sampler_traffic.json
if len(sys.argv) > 2 and len(rows) > 2:
    import json
    r = rows[2]

    def val(key):
        i = hdr.index(key)
        v, u = float(r[i]), units[i].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)

    json.dump({"kernel": r[hdr.index("Kernel Name")][:120], "dram_bytes_read": val("dram__bytes_read.sum"),
               "dram_bytes_write": val("dram__bytes_write.sum"), "duration_ms_under_ncu": float(r[hdr.index("gpu__time_duration.sum")]),
               "source": f"ncu --set full capture {rep} (one launch of the sampler of bench.py's C4 step), summarised by "
                         f"tools/ncu_summary.py into profiles/"},
              open(sys.argv[2], "w"), indent=1)
