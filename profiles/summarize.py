"""Turn the ncu reports captured by profiles/capture.sh (gpurun_out/*.ncu-rep, read here with `ncu -i`) into the
markdown tables committed under profiles/.  Usage: python profiles/summarize.py"""
import csv, io, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"), ("gpu__time_duration.sum", "time us"),
        ("dram__bytes_read.sum", "DRAM rd MB"), ("dram__bytes_write.sum", "DRAM wr MB"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM GB"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe %"),
        ("sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "uniform pipe %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %")]


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def num(v, unit):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    scale = {"Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "byte": 1e-6, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit)
    if scale is not None and unit.endswith("byte"):
        return f"{x * scale:.1f}"
    if scale is not None:
        return f"{x * scale:.1f}"
    return f"{x:.1f}" if abs(x) < 1e6 else f"{x:.3g}"


def table(rep):
    hdr, units, rows = rows_of(rep)
    ki = hdr.index("Kernel Name")
    cols = [(hdr.index(m), label, m) for m, label in WANT if m in hdr]
    lines = ["| kernel | " + " | ".join(l for _, l, _ in cols) + " |", "|---|" + "---|" * len(cols)]
    for r in rows:
        name = r[ki].replace("void unnamed>::", "").split("(")[0]
        vals = []
        for i, label, m in cols:
            v = num(r[i], units[i])
            if label == "L2->SM GB":
                v = f"{float(v) / 1e3:.3f}" if units[i] == "Mbyte" else v
            vals.append(v)
        lines.append(f"| `{name}` | " + " | ".join(vals) + " |")
    return "\n".join(lines)


def launches(path):
    """share table of an `ncu --metrics gpu__time_duration.sum --csv` launch list"""
    import re
    agg = {}
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("jck::<unnamed>::", "")
        name = re.sub(r"at::native::|at::<unnamed>::|<unnamed>::", "", name)[:90]
        v = float(r["Metric Value"].replace(",", ""))
        us = v / 1e3 if r["Metric Unit"] in ("ns", "nsecond") else v
        t = agg.setdefault(name, [0.0, 0])
        t[0] += us
        t[1] += 1
    total = sum(t[0] for t in agg.values())
    out = [f"{sum(t[1] for t in agg.values())} launches, {total / 1e3:.2f} ms serialised", "",
           "| share | total us | launches | avg us | kernel |", "|---:|---:|---:|---:|---|"]
    for name, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        out.append(f"| {100 * us / total:.1f}% | {us:.0f} | {n} | {us / n:.1f} | `{name}` |")
    return "\n".join(out)


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        print(launches(sys.argv[2]))
        sys.exit(0)
    for rep in sys.argv[1:]:
        print(table(rep))
        print()
