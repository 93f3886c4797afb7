"""Per-kernel totals of ONE steady-state step from an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py.
A step starts at the first kernel of the forward pass (the padded channels-last copy of the input batch); the last complete
step of the capture is summarised.  python tools/launch_summary.py gpurun_out/launches.csv [out.csv] > profiles/x.md"""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = name.replace("mas::", "").replace("void ", "").replace("(anonymous namespace)::", "")
    name = re.sub(r"\(bool\)", "", name)
    name = re.sub(r"\((?:const |unsigned |float|double|int|long|void|__half|uint|CUtensorMap|mas_|tc::|HParams|WTParams|WParams|Params|P3).*$", "", name)
    return name.strip()


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
    starts = [i for i, r in enumerate(rows) if "nchw_to_nhwc_pad" in r[4]]
    if len(starts) < 2:
        raise SystemExit("need at least two step starts in the capture")
    a, b = starts[-2], starts[-1]
    step = rows[a:b]
    agg = OrderedDict()
    for r in step:
        k = short(r[4])
        c, t = agg.get(k, (0, 0.0))
        agg[k] = (c + 1, t + float(r[14]) / 1e6)
    tot = sum(t for _, t in agg.values())
    print("# One steady-state training step (batch 32, 256x256) under `ncu --metrics gpu__time_duration.sum --clock-control none`")
    print("# (bench.py --steps 1 --warmup 1 --step-only; per-launch times are serialised / cold-cache: the SHARES are what to compare")
    print("# with bench.py --profile).  %d launches, %.2f ms of kernel time.\n" % (len(step), tot))
    print("| kernel | launches | ms | share |\n|---|---|---|---|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.1f %% |" % (k, c, t, 100 * t / tot))
    if len(sys.argv) > 2:
        with open(sys.argv[2], "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["index", "kernel", "grid", "block", "ns"])
            for i, r in enumerate(step):
                w.writerow([i, short(r[4]), r[8], r[7], r[14]])


if __name__ == "__main__":
    main()
