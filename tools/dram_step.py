"""Sum DRAM read+write bytes of every launch of one step from an ncu CSV (run anywhere):

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/dram_step.csv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-gpu-eager
    python tools/dram_step.py gpurun_out/dram_step.csv LAUNCHES_PER_STEP ave 160 > profiles/r02_dram_step.json

The launch list holds warm-up + timed + profiling steps of the bench; the LAST `LAUNCHES_PER_STEP` launches of this library
(kernel names in namespace avmoe) plus the torch kernels interleaved with them are taken as one steady-state step."""
import collections, csv, json, re, sys


def main(path, launches_per_step, variant, BT):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        i = int(row["ID"])
        d = per.setdefault(i, {"name": row["Kernel Name"], "rd": 0.0, "wr": 0.0, "ns": 0.0})
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        m = row["Metric Name"]
        if m.startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
            d["rd" if "read" in m else "wr"] += v
        elif m == "gpu__time_duration.sum":
            d["ns"] += v * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(u, 1.0)
    ids = list(per)
    ours = [i for i in ids if "avmoe::" in per[i]["name"]]
    first = ours[-launches_per_step] if launches_per_step and len(ours) >= launches_per_step else ids[0]
    sel = [per[i] for i in ids if i >= first]
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for d in sel:
        k = re.sub(r"\(.*", "", d["name"]).replace("void ", "")[:60]
        agg[k][0] += 1; agg[k][1] += d["rd"] + d["wr"]; agg[k][2] += d["ns"] / 1e6
    out = {"variant": variant, "BT": BT, "launches": len(sel), "dram_bytes": sum(d["rd"] + d["wr"] for d in sel),
           "dram_read": sum(d["rd"] for d in sel), "dram_write": sum(d["wr"] for d in sel), "serialised_ms": sum(d["ns"] for d in sel) / 1e6,
           "by_kernel": {k: {"n": v[0], "GB": round(v[1] / 1e9, 3), "ms": round(v[2], 3)} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]}}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), sys.argv[3], int(sys.argv[4]))
