"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel share table (markdown)."""
import collections
import csv
import sys


def main(path, title, top=25):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ix = {h: i for i, h in enumerate(hdr)}
    tot = collections.defaultdict(lambda: [0, 0.0])
    for row in r:
        if len(row) < len(hdr) or row[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(row[ix["Metric Value"]].replace(",", ""))
        u = row[ix["Metric Unit"]]
        v = v / 1000 if u in ("nsecond", "ns") else v * 1000 if u in ("msecond", "ms") else v
        t = tot[row[ix["Kernel Name"]]]
        t[0] += 1
        t[1] += v
    T = sum(v[1] for v in tot.values())
    print(f"# {title}\n")
    print(f"Total device time in the list: {T / 1000:.1f} ms over {sum(v[0] for v in tot.values())} launches "
          "(cold-cache, serialised under ncu: compare SHARES, not absolutes).\n")
    print("| share | total us | launches | us/launch | kernel |\n|---|---|---|---|---|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"| {100 * v[1] / T:.1f}% | {v[1]:.0f} | {v[0]} | {v[1] / v[0]:.1f} | `{k[:100]}` |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
