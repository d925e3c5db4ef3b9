"""Aggregate an ncu --csv launch list (dram bytes, time, instructions) per kernel name."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
iN, iM, iV = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
iU = hdr.index("Metric Unit")
iID = hdr.index("ID")
agg = collections.defaultdict(lambda: collections.defaultdict(float))
cnt = collections.Counter()
seen = set()
for r in rows[1:]:
    name = r[iN].split("(")[0].split("::")[-1]
    v = float(r[iV].replace(",", ""))
    u = r[iU]
    if u == "Mbyte": v *= 1e6
    elif u == "Kbyte": v *= 1e3
    elif u == "Gbyte": v *= 1e9
    elif u in ("us", "usecond"): v *= 1e-3
    elif u in ("ns", "nsecond"): v *= 1e-6
    elif u in ("ms", "msecond"): pass
    agg[name][r[iM]] += v
    if (r[iID]) not in seen:
        seen.add(r[iID]); cnt[name] += 1
tot = collections.defaultdict(float)
print("%-22s %4s %10s %10s %10s %12s" % ("kernel", "n", "time_ms", "rd_MB", "wr_MB", "inst_M"))
for name, m in sorted(agg.items(), key=lambda kv: -kv[1].get("gpu__time_duration.sum", 0)):
    t, rd, wr, ins = m.get("gpu__time_duration.sum", 0), m.get("dram__bytes_read.sum", 0), m.get("dram__bytes_write.sum", 0), m.get("smsp__inst_executed.sum", 0)
    print("%-22s %4d %10.4f %10.2f %10.2f %12.2f" % (name, cnt[name], t, rd / 1e6, wr / 1e6, ins / 1e6))
    for k, v in (("t", t), ("rd", rd), ("wr", wr), ("ins", ins)): tot[k] += v
print("%-22s %4d %10.4f %10.2f %10.2f %12.2f" % ("TOTAL", sum(cnt.values()), tot["t"], tot["rd"] / 1e6, tot["wr"] / 1e6, tot["ins"] / 1e6))
if len(sys.argv) > 2:  # json for bench.py's roofline.traffic: per kernel name of one pair
    import json
    per = {}
    for name, m in sorted(agg.items(), key=lambda kv: -kv[1].get("gpu__time_duration.sum", 0)):
        per[name] = {"launches": cnt[name], "time_ms": round(m.get("gpu__time_duration.sum", 0), 4),
                     "dram_read_MB": round(m.get("dram__bytes_read.sum", 0) / 1e6, 2),
                     "dram_write_MB": round(m.get("dram__bytes_write.sum", 0) / 1e6, 2),
                     "warp_inst_M": round(m.get("smsp__inst_executed.sum", 0) / 1e6, 2)}
    json.dump({"what": sys.argv[3] if len(sys.argv) > 3 else "", "per_pair": per}, open(sys.argv[2], "w"), indent=1)
