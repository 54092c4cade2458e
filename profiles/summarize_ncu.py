"""Turn an `ncu --set full` report into the markdown table committed under profiles/ (run where ncu is installed,
no GPU needed):   python profiles/summarize_ncu.py <report.ncu-rep | raw.csv> "title" [--json traffic.json]
One row per launch: duration, DRAM bytes, DRAM %, tensor-pipe %, L2 / L1 throughput %, registers, instructions."""
import csv
import io
import json
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dram rd MB"), ("dram__bytes_write.sum", "dram wr MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM fill"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "instructions")]


def main():
    rep, title = sys.argv[1], sys.argv[2]
    # a report, or the `ncu -i report --page raw --csv` text of one (what the GPU job brings back: reports are too big)
    raw = open(rep).read() if rep.endswith(".csv") else \
        subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    cols = [(hdr.index(m), lab, units[hdr.index(m)]) for m, lab in WANT if m in hdr]
    ik, ig = hdr.index("Kernel Name"), hdr.index("Grid Size")
    print(f"# {title}\n")
    print("| kernel | grid | " + " | ".join(f"{lab} ({u})" if u and u not in lab else lab for _, lab, u in cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    tot_rd = tot_wr = n = 0
    for r in body:
        name = r[ik].split("(")[0].replace("void ", "").replace("nrl::", "")
        vals = []
        for i, lab, u in cols:
            try:
                v = float(r[i].replace(",", ""))
                vals.append(f"{v:.1f}" if abs(v) < 1e5 else f"{v:.3g}")
            except ValueError:
                vals.append(r[i])
        print(f"| {name[:44]} | {r[ig]} | " + " | ".join(vals) + " |")
        try:
            scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
            ird, iwr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            tot_rd += float(r[ird]) * scale.get(units[ird], 1.0)
            tot_wr += float(r[iwr]) * scale.get(units[iwr], 1.0)
            n += 1
        except (ValueError, KeyError):
            pass
    if "--json" in sys.argv and n:
        out = {"source": rep, "launches": n, "dram_bytes_read_total": tot_rd, "dram_bytes_write_total": tot_wr,
               "dram_bytes_per_launch_avg": (tot_rd + tot_wr) / n}
        json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
