"""Turn ncu captures brought back in gpurun_out/ into the tracked summaries under profiles/.
usage: python tools/summarize_profiles.py <tag> <full.ncu-rep> <launches.csv>"""
import collections, csv, subprocess, sys

tag, rep, launches = sys.argv[1:4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keep = ['gpu__time_duration.sum', 'Grid Size', 'Block Size', 'launch__registers_per_thread', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio']
with open(f"profiles/{tag}_ncu_full.md", "w") as f:
    f.write(f"# {tag}: `ncu --set full --clock-control none --import-source on -k regex:\"dense_score|paths_topk\" -s 6 -c 2 "
            "python bench.py --steps 2 --warmup 1 --no-cpu-baseline`\n\n"
            "cfg3 (125,659 nodes x 768-d, 10,000 queries per launch).  Values per launch.\n\n")
    for r in rows[2:]:
        f.write(f"## {r[hdr.index('Kernel Name')].split('(')[0]}\n\n| metric | value | unit |\n|---|---|---|\n")
        for k in keep:
            if k in hdr:
                i = hdr.index(k)
                f.write(f"| {k} | {r[i]} | {units[i]} |\n")
        f.write("\n")
lr = list(csv.reader(l for l in open(launches) if l.startswith('"')))
h = lr[0]
ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
agg = collections.OrderedDict()
for r in lr[1:]:
    name = r[ki].split('(')[0][:60]
    v = float(r[vi].replace(',', ''))
    v = {'ns': v / 1e6, 'us': v / 1e3, 'ms': v, 's': v * 1e3}[r[ui]]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
with open(f"profiles/{tag}_launches.md", "w") as f:
    f.write(f"# {tag}: `ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 2 --warmup 1 "
            "--no-cpu-baseline`\n\nPer-launch times under ncu are cold-cache and serialised: compare shares.  The list covers the "
            "whole process (setup ifit, index build, warm-up, timed steps, kernel-only loop, best-first sample).\n\n"
            "| kernel | launches | total ms | share |\n|---|---|---|---|\n")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if ms / tot > 0.0001:
            f.write(f"| {k} | {n} | {ms:.2f} | {ms / tot * 100:.2f}% |\n")
    step = {k: v for k, v in agg.items() if any(s in k for s in ['dense_score', 'paths_topk', 'merge_topk', 'tile_queries'])}
    st = sum(v[1] for v in step.values())
    f.write("\nWithin the predict step:\n\n| kernel | share of step |\n|---|---|\n")
    for k, (n, ms) in sorted(step.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| {k} | {ms / st * 100:.1f}% |\n")
print(open(f"profiles/{tag}_launches.md").read()[-700:])
