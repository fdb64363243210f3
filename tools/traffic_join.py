"""Join an `ncu --set full` raw CSV of one adapter call's GEMM launches with the algorithmic bytes of the same launches
(tools/call_gemm_bytes.py) -> profiles/r01_traffic_gemm.json (read by bench.py for roofline.traffic).  Runs anywhere.
usage: traffic_join.py <ncu_raw.csv> <call_gemm_bytes.json> [ridge_flop_per_byte]"""
import csv, json, sys


def main(raw_csv, bytes_json, ridge=207.5, out="profiles/r01_traffic_gemm.json", summary="profiles/r01_ncu_full_gemm_ave_stage1_visual.csv"):
    launches = json.load(open(bytes_json))
    rows = list(csv.reader(open(raw_csv)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, units, data = rows[hi], rows[hi + 1], [r for r in rows[hi + 2:] if r]
    col = {n: i for i, n in enumerate(hdr)}
    assert len(launches) == len(data), (len(launches), len(data))
    want = [w for w in ("Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
                        "launch__registers_per_thread", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct") if w in col]
    with open(summary, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["gemm_shape"] + want); w.writerow([""] + [units[col[x]] for x in want])
        for g, r in zip(launches, data):
            w.writerow([g["tag"]] + [r[col[x]] for x in want])
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
    per, tot_alg, tot_dram, n = [], 0.0, 0.0, 0
    for g, r in zip(launches, data):
        dram = float(r[col["dram__bytes_read.sum"]]) * scale[units[col["dram__bytes_read.sum"]]] + \
            float(r[col["dram__bytes_write.sum"]]) * scale[units[col["dram__bytes_write.sum"]]]
        hb = g["flops"] / g["bytes"] < ridge
        per.append(dict(tag=g["tag"], hbm_bound=hb, algorithmic_MB=round(g["bytes"] / 1e6, 1), dram_MB=round(dram / 1e6, 1),
                        ncu_us=float(r[col["gpu__time_duration.sum"]]), event_us=round(g["ms"] * 1e3, 1)))
        if hb:
            tot_alg += g["bytes"]; tot_dram += dram; n += 1
    json.dump({"source": "ncu --set full --clock-control none on tools/profile_one.py ave 1 visual 160 (bf16), 2nd iteration; algorithmic bytes "
                         "from tools/call_gemm_bytes.py on the same call; joined by tools/traffic_join.py",
               "hbm_bound_gemm_launches": n, "algorithmic_MB": round(tot_alg / 1e6, 1), "dram_read_plus_write_MB": round(tot_dram / 1e6, 1),
               "dram_to_algorithmic_ratio": round(tot_dram / tot_alg, 4), "per_launch": per}, open(out, "w"), indent=1)
    print(n, "HBM-bound launches; dram/algorithmic =", round(tot_dram / tot_alg, 4))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 207.5)
