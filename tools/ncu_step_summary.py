"""Summarise the raw ncu CSV of one step (tools/profile_step.py under `ncu --csv --metrics ...`) into the
per-kernel table committed under profiles/ and the DRAM traffic of the two kernel families."""
import collections
import csv
import json
import re
import sys

raw, out_csv, out_json = sys.argv[1:4]
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
    v = float(r[ix["Metric Value"]].replace(",", ""))
    a = agg.setdefault(k, collections.defaultdict(float))
    if m == "gpu__time_duration.sum":
        a["n"] += 1
        a["us"] += v / 1000 if u == "ns" else (v if u == "us" else v * 1000)
        a["_t"] = v / 1000 if u == "ns" else (v if u == "us" else v * 1000)
    elif m.startswith("dram__bytes"):
        a["dram"] += v * mult[u]
    elif m.startswith("lts__t_bytes"):
        a["l2"] += v * mult[u]
    elif m.startswith("sm__pipe_tensor"):
        a["tensor_w"] += v * a["_t"]
    elif m.startswith("smsp__issue_active"):
        a["issue_w"] += v * a["_t"]
tot = sum(a["us"] for a in agg.values())
with open(out_csv, "w") as f:
    f.write("# ncu per-kernel metrics of ONE fusion-block step (B=16, dropout 0.1, tf32); times are cold-cache and "
            "serialised (compare SHARES). total %.1f us over %d launches\n" % (tot, sum(a["n"] for a in agg.values())))
    f.write("kernel,launches,total_us,share,dram_MB_per_launch,l2_MB_per_launch,tensor_pipe_pct(time-weighted),"
            "issue_active_pct(time-weighted)\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        f.write("%s,%d,%.1f,%.3f,%.1f,%.1f,%.1f,%.1f\n" % (k, a["n"], a["us"], a["us"] / tot, a["dram"] / a["n"] / 1e6,
                                                          a["l2"] / a["n"] / 1e6, a["tensor_w"] / a["us"], a["issue_w"] / a["us"]))
fam = {}
for name, pat in (("gemm", "gemm_tf32"), ("attention", "attn")):
    sel = [a for k, a in agg.items() if pat in k]
    n = sum(a["n"] for a in sel)
    fam[name] = {"launches": int(n), "dram_bytes_per_launch": sum(a["dram"] for a in sel) / max(n, 1),
                 "total_dram_bytes": sum(a["dram"] for a in sel), "total_us_under_ncu": sum(a["us"] for a in sel)}
fam["source"] = "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, " + out_csv
json.dump(fam, open(out_json, "w"), indent=1)
print(open(out_csv).read())
