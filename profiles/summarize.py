"""Turn ncu outputs brought back in gpurun_out/ into the compact summaries committed here.
  python profiles/summarize.py launches gpurun_out/launches.csv  > profiles/rNN_launches.md
  python profiles/summarize.py full gpurun_out/prof.ncu-rep      > profiles/rNN_ncu_full.md
  python profiles/summarize.py traffic gpurun_out/prof.ncu-rep 32 "<what was captured>" > profiles/ncu_traffic.json
"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg, order = {}, []
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        name = r[ik].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        t = float(r[iv].replace(",", ""))
        if name not in agg:
            agg[name] = [0.0, 0]
            order.append(name)
        agg[name][0] += t
        agg[name][1] += 1
    unit = rows[1][hdr.index("Metric Unit")] if len(rows) > 1 else "?"
    tot = sum(v[0] for v in agg.values())
    print("| kernel | launches | total (%s) | share |" % unit)
    print("|---|---:|---:|---:|")
    for n in sorted(agg, key=lambda k: -agg[k][0]):
        print("| `%s` | %d | %.1f | %.1f%% |" % (n[:90], agg[n][1], agg[n][0], 100 * agg[n][0] / tot))
    print("\ntotal over captured launches: %.1f %s (cold-cache, serialised: compare shares)" % (tot, unit))


NAME_MAP = [("pathnet_fwd_kernel", "pathnet_fwd"), ("pathnet_dw_kernel", "pathnet_dw"), ("pool_bwd_dz_kernel", "pool_bwd_dz"),
            ("pool_tc_kernel<32, 0>", "pool_fwd"), ("pool_tc_kernel<64, 0>", "pool_fwd"),
            ("pool_tc_kernel<32, 1>", "pool_bwd_dq"), ("pool_tc_kernel<64, 1>", "pool_bwd_dq"),
            ("modularity_sweep_kernel", "modularity_sweep"), ("modularity_prep_tc_kernel", "modularity_prep"),
            ("modularity_degrees_closed_kernel", "modularity_degrees_closed"), ("modularity_finish_tc_kernel", "modularity_finish")]


def traffic(path, bags, source):
    """profiles/ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch, keyed by the launch names
    bench.py uses (mean over the captured launches of a kernel)."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    acc = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        key = next((v for k, v in NAME_MAP if k in name), None)
        if key is None:
            continue
        b = sum(float(r[idx[m]].replace(",", "")) * scale[units[idx[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        acc.setdefault(key, []).append(b)
    print(json.dumps({"bags_per_step": int(bags), "source": source,
                      "bytes_per_launch": {k: sum(v) / len(v) for k, v in sorted(acc.items())}}, indent=1))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("### `%s`\n" % r[idx["Kernel Name"]][:110])
        print("| metric | value | unit |\n|---|---:|---|")
        for k in KEYS:
            if k in idx:
                print("| %s | %s | %s |" % (k, r[idx[k]], units[idx[k]]))
        print()


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
