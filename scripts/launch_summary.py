"""Summarise the ncu launch list of the default bench command (scripts/profile_round.sh) into profiles/: per-kernel totals and the
share of dpr_kernel inside a step (a step = dpr_prep + dpr_kernel [fused K1+K4] + any_flag + masked pyramid + dpr_prep +
masked dpr_kernel), to be compared with the CUDA-event share bench.py reports (kernel_ms / ms_per_step).

    python scripts/launch_summary.py gpurun_out/r02b_bench_launches.csv gpurun_out/r02b_bench_launches_plain.json profiles/r02_bench_launches_summary.txt
"""
import collections, csv, json, sys


def main():
    src, bench, dst = sys.argv[1:4]
    rows = list(csv.DictReader([l for l in open(src) if not l.startswith("==")]))
    def us(x):
        v = float(x["Metric Value"].replace(",", "")); u = x["Metric Unit"]
        return v / 1e3 if u.startswith("ns") else (v if u.startswith("us") else v * 1e3)
    names = [r["Kernel Name"] for r in rows]
    t = [us(r) for r in rows]
    out = ["# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 of: python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-streams",
           "# cold-cache, serialised launch times: compare SHARES, not absolutes", ""]
    # steps: main dpr_kernel launches are the ones longer than 100 us
    steps = []
    for i, n in enumerate(names):
        if "dpr_kernel" in n and t[i] > 100.0:
            j = i - 1                      # its dpr_prep
            k = i + 1
            while k < len(names) and not ("dpr_kernel" in names[k] and t[k] > 100.0) and ("dpr_" in names[k] or "any_flag" in names[k] or "pyr_down_stream_kernel" in names[k] and t[k] < 100.0):
                k += 1
            steps.append((t[i], sum(t[j:k])))
    line = json.loads([l for l in open(bench) if l.startswith("{")][-1])
    steps = steps[: line["warmup"] + line["steps"]]          # what follows are the post-run comparisons (K4 alone on a built pyramid)
    out.append(f"{'step':>4s} {'dpr_kernel us':>14s} {'whole step us':>14s} {'share':>8s}")
    for s, (a, b) in enumerate(steps):
        out.append(f"{s:4d} {a:14.1f} {b:14.1f} {100 * a / b:7.2f}%")
    ev = line["kernel_ms"]["dense_refinement_with_fused_pyramid"] / line["ms_per_step"]
    out.append(f"share of dpr_kernel in a step: {100 * sum(a for a, _ in steps) / sum(b for _, b in steps):.2f} % under ncu, "
               f"{100 * ev:.2f} % by CUDA events in the plain run of the same command ({line['kernel_ms']['dense_refinement_with_fused_pyramid']:.3f} of {line['ms_per_step']:.3f} ms)")
    out.append("")
    agg = collections.OrderedDict()
    for n, v in zip(names, t):
        a = agg.setdefault(n[:86], [0, 0.0]); a[0] += 1; a[1] += v
    out.append(f"{'kernel (whole run incl. frame generation and the post-run comparisons)':88s} {'launches':>8s} {'total us':>12s} {'avg us':>10s}")
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{n:88s} {c:8d} {v:12.1f} {v / c:10.2f}")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out[:14]))


if __name__ == "__main__":
    main()
