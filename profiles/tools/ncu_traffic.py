"""ncu raw CSV of profiles/tools/sa_branch.py (one SA1 branch, P = 2,097,152 rows) -> profiles/ncu_traffic_r02.json:
{"<entry point>|<key fields of bench.py's kernel table>": dram__bytes_read.sum + dram__bytes_write.sum per launch}.
Usage: ncu -i X.ncu-rep --page raw --csv > raw.csv; python profiles/tools/ncu_traffic.py raw.csv [summary.txt]"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
def val(r, name):
    v, u = float(r[col[name]]), units[col[name]]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
MAP = [  # (substrings of the kernel name, key in bench.py's table); generation 4 (rowgemm_ws2_kernel) serves these shapes
    (("rowgemm_ws2_kernel", "WProGatherBnAct", "WEpiStoreStats"), "pcl_rowgemm|sa_l2|2|1|2097152|64|96"),
    (("rowgemm_ws2_kernel", "WProBnAct", "WEpiMaxMinStats"), "pcl_rowgemm|sa_l3|1|2|2097152|96|128"),
    (("rowgemm_ws2_kernel", "WProBnAct", "WEpiBwdYMaskRouted"), "pcl_rowgemm|sa_b3|1|7|2097152|96|96"),
    (("rowgemm_ws_kernel", "WProG3A2", "WEpiBwdYMask"), "pcl_rowgemm|sa_b3|4|6|2097152|224|96"),
    (("rowgemm_ws2_kernel", "WProBnBwd", "WEpiStore"), "pcl_rowgemm|sa_b2|3|0|2097152|96|64"),
    (("wgrad_ws_kernel", "GBnAct, ws::GBnActOnes"), "pcl_wgrad|sa_gram|2097152|96|97"),
    (("wgrad_own_kernel", "GGatherBnActMask"), "pcl_wgrad|sa_dw2|2097152|96|128"),
    (("wgrad_ws_kernel", "GBnBwd", "GGatherBnActMask"), "pcl_wgrad|sa_dw2_ws|2097152|96|128"),
    (("sel_outer_group_kernel",), "pcl_sel_outer|sa_sel_outer|16384|128|96"),
    (("gather_bn_backward_kernel<1>",), "pcl_gather_bn_backward_masked|sa_b1_scatter|2097152|64"),
    (("gather_stats_kernel",), "pcl_gather_stats|sa_gather_stats|2097152|64"),
    (("routed_sort_kernel",), "pcl_routed_sort|sa_routed_sort|16384|128"),
]
out, lines = {}, []
want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "lts__t_sector_hit_rate.pct"]
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    t = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    key = next((k for subs, k in MAP if all(s in name for s in subs)), None)
    if key and key not in out:
        out[key] = int(t)
    lines.append(f"---- {name[:150]}\n  bench key: {key}\n  dram read+write: {t/1e6:.1f} MB (read {val(r,'dram__bytes_read.sum')/1e6:.1f}, write {val(r,'dram__bytes_write.sum')/1e6:.1f})")
    for w in want:
        if w in col:
            lines.append(f"  {w}: {r[col[w]]} {units[col[w]]}")
try:   # merge into the existing table (captures of single kernels update their rows only)
    prev = json.load(open("profiles/ncu_traffic_r02.json"))
except Exception:
    prev = {}
prev.update(out)
out = prev
json.dump(out, open("profiles/ncu_traffic_r02.json", "w"), indent=1)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write("\n".join(lines) + "\n")
print(json.dumps(out, indent=1))
