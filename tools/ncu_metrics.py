#!/usr/bin/env python3
"""Extract the metrics quoted in profiles/*.md from an ncu report (`ncu -i rep --page raw --csv`) into JSON.
usage: ncu_metrics.py <report.ncu-rep> <out.json> [traffic.json]"""
import csv, json, re, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = re.compile(r"^(dram__bytes_(read|write)\.sum$|gpu__dram_throughput\.avg\.pct|gpu__time_duration\.sum|launch__(block_size|grid_size|registers_per_thread|shared_mem_per_block_dynamic)$|"
                  r"sass__inst_executed_register_spilling$|sm__icc_request_hit_rate|sm__inst_executed\.avg\.per_cycle_active|sm__inst_executed\.sum$|smsp__inst_executed\.sum$|"
                  r"sm__inst_executed_pipe_(alu|fma|lsu|xu)\.avg\.pct_of_peak_sustained_active|sm__pipe_tensor_cycles_active|sm__warps_active\.avg\.pct|"
                  r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio|smsp__thread_inst_executed_per_inst_executed\.ratio|smsp__issue_active\.avg\.pct|"
                  r"l1tex__t_sector_hit_rate\.pct|lts__t_sector_hit_rate\.pct|smsp__inst_executed_op_local|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$)")
d = {h: [v, u] for h, u, v in zip(hdr, units, vals) if keep.search(h)}
d["Kernel Name"] = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
json.dump(d, open(out, "w"), indent=1, sort_keys=True)
def num(k):
    v, u = d[k]; x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
if len(sys.argv) > 3:
    r, w = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
    json.dump({"dram_bytes_per_launch": int(r + w), "source": f"ncu --set full, one steady-state launch of bench.py: dram__bytes_read.sum ({int(r):,}) + dram__bytes_write.sum ({int(w):,})",
               "report": rep + " (scratch)"}, open(sys.argv[3], "w"))
print(json.dumps({k: d[k] for k in sorted(d) if "stalled" not in k}, indent=1))
