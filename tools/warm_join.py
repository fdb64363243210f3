"""Join an ncu launch list of tools/profile_one.py (last iteration) with the GEMM shape tags it dumped."""
import csv, json, re, sys, collections

def main(csv_path, tags_path):
    lines = [l for l in open(csv_path) if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    tags = json.load(open(tags_path))
    n = len(rows) // 2                      # two identical iterations; keep the second
    rows = rows[len(rows) - n:]
    # the second iteration launches a few more kernels than the first (gradient accumulation), so the cut above may drop the first
    # GEMMs of the iteration: align the tags from the END of the list
    n_g = sum(1 for r in rows if "gemm_tcgen05" in r["Kernel Name"])
    gi = 0
    tot = 0.0
    out = []
    for r in rows:
        k = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("avmoe::", "")[:44]
        us = float(r["Metric Value"].replace(",", "")) / 1e3
        tot += us
        if "gemm_tcgen05" in k:
            ti = len(tags) - n_g + gi
            k = "GEMM " + (tags[ti] if 0 <= ti < len(tags) else "?")
            gi += 1
        out.append((k, us, r["Grid Size"]))
    print(f"{len(rows)} launches, {tot:.1f} us serialized (warm L2), {gi} GEMMs / {len(tags)} tags")
    for k, us, g in out:
        print(f"{us:8.1f}  {g:>16s}  {k}")

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
