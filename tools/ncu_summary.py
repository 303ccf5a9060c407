#!/usr/bin/env python3
"""Summarise ncu exports for profiles/: tools/ncu_summary.py <raw.csv> <source.csv> <launches.csv> <tag>"""
import collections, csv, json, sys

raw, src, launches, tag = sys.argv[1:5]
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.max', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum', 'lts__t_bytes.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'smsp__sass_average_data_bytes_per_sector_mem_global_op_st.ratio', 'dram__bytes_write.sum.pct_of_peak_sustained_elapsed']
out = {"kernels": []}
for r in rows[2:]:
    out["kernels"].append({k: (r[hdr.index(k)] + ' ' + units[hdr.index(k)]).strip() for k in want if k in hdr})
rows = list(csv.reader(open(src)))
h = rows[1]
body = []
for r in rows[2:]:
    if len(r) < 6 or r[0] in ('Kernel Name', 'Address'):
        break
    body.append(r)
ix = h.index('Instructions Executed')
total = sum(int(r[ix]) for r in body)
mix = collections.Counter()
for r in body:
    t = r[1].split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    mix[op] += int(r[ix])
out["warp_instructions_executed"] = total
out["dynamic_mix_fraction"] = {k: round(v / total, 4) for k, v in mix.most_common(16)}
L = list(csv.DictReader(l for l in open(launches) if l.startswith('"')))
tot = sum(float(d['Metric Value']) for d in L)
gen = sum(float(d['Metric Value']) for d in L if 'genbod' in d['Kernel Name'])
out["launch_list"] = {"launches": len(L), "genbod_share_of_gpu_time": gen / tot,
                      "genbod_times_ns": [float(d['Metric Value']) for d in L if 'genbod' in d['Kernel Name']]}
json.dump(out, open(f"profiles/{tag}_ncu_summary.json", "w"), indent=1)
with open(f"profiles/{tag}_launches.csv", "w") as f:
    f.write("id,kernel,grid,block,gpu_time_ns\n")
    for d in L:
        f.write(f"{d['ID']},\"{d['Kernel Name'][:100]}\",\"{d['Grid Size']}\",\"{d['Block Size']}\",{d['Metric Value']}\n")
print(json.dumps(out, indent=1)[:3500])
