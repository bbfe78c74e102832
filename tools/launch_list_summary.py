#!/usr/bin/env python3
"""ncu launch list (gpu__time_duration.sum per launch, csv) -> markdown table of kernel shares.
usage: python tools/launch_list_summary.py gpurun_out/launches_rX.csv profiles/launches_rX.md "<command line>" """
import csv, io, re, sys, collections
src, out, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
text = open(src).read()
start = text.index('"ID"')
rows = list(csv.DictReader(io.StringIO(text[start:])))
tot = collections.OrderedDict()
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("sw4::", "")
    ns = float(r["Metric Value"]) * {"ns": 1, "us": 1e3, "ms": 1e6}.get(r["Metric Unit"], 1)
    n, t = tot.get(name, (0, 0.0))
    tot[name] = (n + 1, t + ns)
allns = sum(t for _, t in tot.values())
with open(out, "w") as f:
    f.write(f"# ncu launch list of `{cmd}`\n# ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are cold-cache and serialised:\n"
            "# compare SHARES, not absolutes.\n\n| kernel | launches | total ms | share |\n|---|---|---|---|\n")
    for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{name}` | {n} | {t/1e6:.3f} | {100*t/allns:.2f} % |\n")
    f.write(f"| **all** | {sum(n for n, _ in tot.values())} | {allns/1e6:.3f} | 100 % |\n")
print(open(out).read())
