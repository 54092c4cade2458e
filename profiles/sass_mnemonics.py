"""Regenerate profiles/r02_sass_mnemonics.md: per-kernel counts of the SASS instructions that prove which hardware path a
kernel of the shipped library uses (`cuobjdump -sass newsreclib_b200/libnrl_b200.so`, run in the build container).
Usage: python profiles/sass_mnemonics.py"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "newsreclib_b200", "libnrl_b200.so")
COLS = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "HMMA", "LDSM", "MOVM"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    counts, name = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            counts[name] = collections.Counter()
            continue
        if name is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for c in COLS:
                if op == c or op.startswith(c + "."):
                    counts[name][c] += 1
    names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    rows = {}
    for mangled, nice in zip(counts, names):
        nice = re.sub(r"^void ", "", nice)
        nice = re.sub(r"\(.*$", "", nice).replace("nrl::", "")
        if sum(counts[mangled].values()):
            rows[nice] = counts[mangled]
    out = ["# SASS mnemonics of the shipped library (`cuobjdump -sass newsreclib_b200/libnrl_b200.so`, END of round 2)", "",
           "Counts of the instructions that prove which hardware path a kernel uses: `UTCHMMA` = tcgen05.mma, `LDTM` = "
           "tcgen05.ld (TMEM), `UTMALDG` / `UTMASTG` / `UTMAREDG` = TMA tensor load / store / reduce, `UBLKCP` = "
           "cp.async.bulk (1-D bulk copy), `HMMA` = warp-level mma.sync, `LDSM` = ldmatrix, `MOVM` = movmatrix.  Kernels "
           "with none of them (the SIMT kernels: gather, pooling backward, scatter, Adam, scorer, losses, metrics, exchange) "
           "are left out.  Regenerate with `python profiles/sass_mnemonics.py`.", "",
           "| kernel | " + " | ".join(COLS) + " |", "|---|" + "---|" * len(COLS)]
    for k in sorted(rows):
        out.append(f"| `{k}` | " + " | ".join(str(rows[k][c]) for c in COLS) + " |")
    tot = collections.Counter()
    for r in rows.values():
        tot.update(r)
    out.append("| **total** | " + " | ".join(str(tot[c]) for c in COLS) + " |")
    open(os.path.join(ROOT, "profiles", "r02_sass_mnemonics.md"), "w").write("\n".join(out) + "\n")
    print(f"{len(rows)} kernels;", dict(tot))


if __name__ == "__main__":
    main()
