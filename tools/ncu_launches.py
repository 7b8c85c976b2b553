"""Summarise an `ncu --csv --metrics gpu__time_duration.sum[,dram__bytes_*]` launch list per
kernel family: count, time share and DRAM traffic.
    python tools/ncu_launches.py launches.csv [out.json]
"""
import collections
import csv
import json
import sys

FAMILIES = ["conv_gemm", "cross_attn", "gn_apply", "gn_silu", "film_kernel", "time_mlp", "ingest_x",
            "ingest_seq", "upsample2x", "cfg_posterior", "cfg_step", "philox", "bias_add", "transpose_kv"]
UNIT = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3,
        "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mn, mu, mv, idc = (h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"),
                       h.index("Metric Value"), h.index("ID"))
per = collections.defaultdict(lambda: collections.defaultdict(float))
launch_ids = collections.defaultdict(set)
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    fam = next((f for f in FAMILIES if f in r[kn]), "other")
    val = float(r[mv].replace(",", "")) * UNIT.get(r[mu], 1.0)
    per[fam][r[mn]] += val
    launch_ids[fam].add(r[idc])
tot = sum(v["gpu__time_duration.sum"] for v in per.values())
out = {}
print(f"{'family':16s} {'launches':>8s} {'time us':>10s} {'share':>7s} {'dram rd MB':>11s} {'dram wr MB':>11s}")
for fam, v in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    t = v["gpu__time_duration.sum"]
    rd, wr = v.get("dram__bytes_read.sum", 0.0), v.get("dram__bytes_write.sum", 0.0)
    n = len(launch_ids[fam])
    out[fam] = {"launches": n, "time_us": t, "share": t / tot, "dram_read_bytes": rd,
                "dram_write_bytes": wr}
    print(f"{fam:16s} {n:8d} {t:10.1f} {100 * t / tot:6.1f}% {rd / 1e6:11.1f} {wr / 1e6:11.1f}")
print(f"{'total':16s} {sum(len(s) for s in launch_ids.values()):8d} {tot:10.1f}")
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)
