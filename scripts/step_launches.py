"""One train step of an `ncu --metrics gpu__time_duration.sum --csv` launch list, in launch order (markdown).

    python scripts/step_launches.py gpurun_out/launches.csv [step_index] > profiles/...md

A step starts at a `take_rows_kernel` pair / the fused `gather_gemm_kernel`; the list is cut at the launches of that kernel."""
import csv
import sys


def main(path, which=3):
    lines = [l for l in open(path) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ix = {h: i for i, h in enumerate(hdr)}
    rows = []
    for row in r:
        if len(row) < len(hdr) or row[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(row[ix["Metric Value"]].replace(",", ""))
        u = row[ix["Metric Unit"]]
        v = v / 1000 if u in ("nsecond", "ns") else v * 1000 if u in ("msecond", "ms") else v
        rows.append((row[ix["Kernel Name"]], v))
    marks = [i for i, (k, _) in enumerate(rows) if "gather_gemm_kernel" in k]
    if len(marks) <= which + 1:
        raise SystemExit(f"only {len(marks)} steps in the list")
    # the step's first launches (index / label lookups) precede the fused gather kernel
    back = 0
    while marks[which] - back - 1 >= 0 and "take_rows" in rows[marks[which] - back - 1][0]:
        back += 1
    a, b = marks[which] - back, marks[which + 1] - back
    step = rows[a:b]
    tot = sum(v for _, v in step)
    ours = sum(1 for k, _ in step if "b200med" in k)
    print(f"Step {which} of the list: {len(step)} launches, {ours} of them b200med kernels, {len(step) - ours} library (torch) launches; "
          f"{tot:.0f} us serialised under ncu (cold caches: compare shares, not absolutes).\n")
    print("| us | share | kernel |\n|---|---|---|")
    for k, v in step:
        print(f"| {v:.1f} | {100 * v / tot:.1f}% | `{k[:120]}` |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3)
