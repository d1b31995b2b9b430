"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--clock-control none --csv` of profiles/one_step.py) into a per-kernel table of the SECOND (warm) step: calls, total
time, DRAM bytes read + written, and the per-step DRAM total the algorithmic figure of SURVEY.md 8(d) is compared with.
Usage: python profiles/launch_summary.py gpurun_out/launches.csv > profiles/rNN_ncu_launches_summary.md"""
import collections, csv, io, sys

ALGO_BYTES_PER_STEP = 4.5e9      # algorithmic HBM bytes of a 512-image step


def load(path):
    txt = open(path).read()
    rd = csv.DictReader(io.StringIO(txt[txt.index('"ID"'):]))
    L = collections.OrderedDict()
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1, 'ms': 1e3, 'usecond': 1,
             'nsecond': 1e-3, 'msecond': 1e3}
    for r in rd:
        d = L.setdefault(int(r['ID']), {'name': r['Kernel Name']})
        d[r['Metric Name']] = float(r['Metric Value'].replace(',', '')) * scale.get(r['Metric Unit'], 1)
    return list(L.values())


def short(n):
    n = n.replace('void ', '').replace('jck::(anonymous namespace)::', '')
    return n.split('(')[0]


def main(path):
    L = load(path)
    starts = [i for i, d in enumerate(L) if 'rng_kernel<true>' in d['name'] or 'rng_kernel<1>' in d['name']]
    step = L[starts[-1]:] if len(starts) >= 2 else L
    agg = collections.OrderedDict()
    for d in step:
        a = agg.setdefault(short(d['name']), [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get('gpu__time_duration.sum', 0.0)
        a[2] += d.get('dram__bytes_read.sum', 0.0)
        a[3] += d.get('dram__bytes_write.sum', 0.0)
    tot_t = sum(a[1] for a in agg.values())
    tot_b = sum(a[2] + a[3] for a in agg.values())
    print(f"Second (warm) eager step of `profiles/one_step.py` (DCGAN 3x64x64, 512 images, bf16) under "
          f"`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none`: "
          f"{len(step)} launches, {tot_t / 1e3:.3f} ms of serialised kernel time (cold caches per launch: shares are "
          f"comparable, absolutes are not), **{tot_b / 1e9:.2f} GB of DRAM traffic per step = {tot_b / ALGO_BYTES_PER_STEP:.2f}x "
          f"the {ALGO_BYTES_PER_STEP / 1e9:.1f} GB algorithmic figure of SURVEY.md 8(d)**.\n")
    print("| kernel | calls | time us | share | DRAM rd MB | DRAM wr MB | GB/s |")
    print("|---|---:|---:|---:|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        gbs = (a[2] + a[3]) / a[1] / 1e3 if a[1] else 0
        print(f"| `{k}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot_t:.1f} % | {a[2] / 1e6:.1f} | {a[3] / 1e6:.1f} | {gbs:.0f} |")


if __name__ == "__main__":
    main(sys.argv[1])
