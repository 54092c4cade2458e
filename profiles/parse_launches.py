"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one steady-state step
(the launches between two consecutive adam_kernel launches).  Usage: parse_launches.py <csv> [title]"""
import collections
import csv
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    seq = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        seq.append((int(row["ID"]), row["Kernel Name"].split("(")[0].replace("void ", ""), v, row["Grid Size"], row["Block Size"]))
    return seq


def one_step(seq):
    """The launches between two consecutive adam_kernel launches; a capture of exactly one profiled step
    (profiles/ncu_step.py: cudaProfilerStart/Stop around one step) is taken whole."""
    ad = [i for i, s in enumerate(seq) if "adam_kernel" in s[1]]
    if len(ad) < 3:
        return seq
    return seq[ad[-3] + 1: ad[-2] + 1]


if __name__ == "__main__":
    seq = load(sys.argv[1])
    step = one_step(seq)
    tot = sum(s[2] for s in step)
    title = sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
    print(f"# {title}\n\n{len(step)} launches, {tot:.0f} us total under ncu (cold-cache, serialised: compare SHARES).\n")
    print("| # | kernel | grid | block | us | share |\n|---|---|---|---|---|---|")
    for s in step:
        print(f"| {s[0]} | {s[1]} | {s[3]} | {s[4]} | {s[2]:.1f} | {100 * s[2] / tot:.1f}% |")
    agg = collections.Counter()
    for s in step:
        agg[s[1]] += s[2]
    print("\n## by kernel\n\n| kernel | us | share |\n|---|---|---|")
    for k, v in agg.most_common():
        print(f"| {k} | {v:.0f} | {100 * v / tot:.1f}% |")
