"""Summarise an `ncu --set full` capture (.ncu-rep, read here without a GPU) into profiles/: a text file with the counters the
roofline discussion uses, and - for the dense-refinement kernel - profiles/ncu_dpr_kernel.json, which bench.py reads for
`roofline.traffic` and `roofline.issue_frac`.

    python scripts/ncu_summary.py gpurun_out/r02_dpr.ncu-rep --kernel dpr_kernel --bench gpurun_out/r02_bench_ncu_dpr.json \
        --out profiles/r02_ncu_dpr_kernel.txt --json profiles/ncu_dpr_kernel.json
"""
import argparse, csv, hashlib, io, json, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum",
    "sm__inst_executed_pipe_xu.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second",
]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]
    return head, units, rows[2:]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--kernel", required=True)
    ap.add_argument("--bench", help="JSON line of the bench run that was profiled (work counts of the captured launch)")
    ap.add_argument("--out", required=True)
    ap.add_argument("--json")
    ap.add_argument("--source-file", default="accurate_aprilgroup_tracking_b200/csrc/agt_dpr.cu")
    a = ap.parse_args()
    head, units, rows = raw_rows(a.rep)
    col = {n: i for i, n in enumerate(head)}
    picked = [r for r in rows if a.kernel in r[col["Kernel Name"]]]
    if not picked:
        sys.exit(f"no launch of {a.kernel} in {a.rep}")
    r = picked[0]
    lines = [f"# {Path(a.rep).name}: first captured launch of {r[col['Kernel Name']][:90]}",
             f"# grid {r[col['Grid Size']]} block {r[col['Block Size']]}; ncu --set full --clock-control none (values under the profiler: cold cache, serialised)"]
    vals = {}
    for name in WANT:
        if name in col:
            v = r[col[name]]
            vals[name] = v
            lines.append(f"{name:95s} {v:>18s} {units[col[name]]}")
    src_sha = hashlib.sha256((ROOT / a.source_file).read_bytes()).hexdigest()
    lines.append(f"# {a.source_file} sha256 {src_sha}")
    Path(a.out).write_text("\n".join(lines) + "\n")
    print("\n".join(lines))
    if a.json:
        f = lambda k: float(vals[k].replace(",", ""))
        bench = {}
        if a.bench:
            bench = json.loads([l for l in Path(a.bench).read_text().splitlines() if l.startswith("{")][-1])
        poses = int(bench.get("config", {}).get("frames_per_gpu", 0))
        dur = f("gpu__time_duration.sum")
        unit = units[col["gpu__time_duration.sum"]]
        ms = dur / 1e6 if unit.startswith("ns") else (dur / 1e3 if unit.startswith("us") else dur)
        js = {"source": Path(a.out).name, "poses": poses, "dram_bytes": f("dram__bytes_read.sum") + f("dram__bytes_write.sum"),
              "warp_instructions": f("smsp__inst_executed.sum"), "sample_evals": bench.get("lm", {}).get("sample_evals"),
              "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"), "kernel_ms_under_ncu": ms,
              "agt_dpr_cu_sha256": src_sha}
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            if units[col[k]].lower().startswith("mbyte"):
                js["dram_bytes"] = (f("dram__bytes_read.sum") + f("dram__bytes_write.sum")) * 1e6
            elif units[col[k]].lower().startswith("gbyte"):
                js["dram_bytes"] = (f("dram__bytes_read.sum") + f("dram__bytes_write.sum")) * 1e9
            elif units[col[k]].lower().startswith("kbyte"):
                js["dram_bytes"] = (f("dram__bytes_read.sum") + f("dram__bytes_write.sum")) * 1e3
        Path(a.json).write_text(json.dumps(js, indent=1) + "\n")
        print(json.dumps(js))


if __name__ == "__main__":
    main()
