"""Summarise the raw ncu CSV of one micro-batch train step (tools/profile_full_step.py under ncu -k regex:corrif) into
a per-kernel table: launches, total time, share, DRAM bytes per launch, tensor-pipe and issue utilisation."""
import collections
import csv
import re
import sys

raw, out_csv = sys.argv[1:3]
rows = list(csv.reader(open(raw)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    k = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("corrif::", "")
    m, u = r[ix["Metric Name"]], r[ix["Metric Unit"]]
    try:
        v = float(r[ix["Metric Value"]].replace(",", ""))
    except ValueError:                      # "n/a" (a metric the kernel does not have)
        continue
    a = agg.setdefault(k, collections.defaultdict(float))
    if m == "gpu__time_duration.sum":
        t = v / 1000 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000)
        a["n"] += 1
        a["us"] += t
        a["_t"] = t
    elif m.startswith("dram__bytes"):
        a["dram"] += v * mult.get(u, 1)
    elif m.startswith("sm__pipe_tensor") or m.startswith("sm__inst_executed_pipe_tensor"):
        a["tensor_w"] += v * a["_t"]
    elif m.startswith("smsp__issue_active"):
        a["issue_w"] += v * a["_t"]
tot = sum(a["us"] for a in agg.values())
with open(out_csv, "w") as f:
    f.write("# ncu per-kernel metrics of the libcorrif_b200 kernels in ONE micro-batch train step (batch 8, 256^2 tiles); times are "
            "cold-cache and serialised (compare SHARES). total %.1f us over %d launches\n" % (tot, sum(a["n"] for a in agg.values())))
    f.write("kernel,launches,total_us,share,dram_MB_per_launch,tensor_pipe_pct(time-weighted),issue_active_pct(time-weighted)\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        f.write("%s,%d,%.1f,%.3f,%.1f,%.1f,%.1f\n" % (k, a["n"], a["us"], a["us"] / tot, a["dram"] / a["n"] / 1e6,
                                                     a["tensor_w"] / max(a["us"], 1e-9), a["issue_w"] / max(a["us"], 1e-9)))
print(open(out_csv).read())
