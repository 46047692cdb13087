"""profiles/r2_traffic.json from an ncu traffic log (tools/ncu_traffic.py's input): per-kernel DRAM bytes of one serial-mode
compression of BASELINE config 2, the source of bench.py's roofline.traffic.   python tools/ncu_traffic_json.py <csv> <out.json>"""
import csv, json, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
d = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) < 15: continue
    d.setdefault((r[0], r[4].split('(')[0]), {})[r[12]] = float(r[14])
names = [("zb_walk_kernel", "cand"), ("zb_parse", "parse_only"), ("zb_merge", "merge"), ("zb_literals", "literals"), ("zb_sequences", "sequences"), ("zb_sizes_scan", "scan"), ("zb_copy", "copy")]
out = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:zb_ ; ZSTDB200_SERIAL=1 python tests/profile_one.py 1024 50 1 1 (datagen -g1GB -P50, level 1, one launch of each kernel over 8192 blocks); raw: " + sys.argv[1],
       "workload": "datagen -g1073741824 -P50, level 1", "kernels": {}}
for (_, kname), v in d.items():
    for pat, key in names:
        if pat in kname:
            k = out["kernels"].setdefault(key, {"dram_read_bytes": 0, "dram_write_bytes": 0, "ncu_duration_ms": 0.0})
            k["dram_read_bytes"] += int(v.get('dram__bytes_read.sum', 0)); k["dram_write_bytes"] += int(v.get('dram__bytes_write.sum', 0))
            k["ncu_duration_ms"] = round(k["ncu_duration_ms"] + v.get('gpu__time_duration.sum', 0) / 1e6, 3)
K = out["kernels"]
def add(a, b):
    return {"dram_read_bytes": a["dram_read_bytes"] + b["dram_read_bytes"], "dram_write_bytes": a["dram_write_bytes"] + b["dram_write_bytes"], "ncu_duration_ms": round(a["ncu_duration_ms"] + b["ncu_duration_ms"], 3)}
if "parse_only" in K and "merge" in K: K["parse"] = add(K["parse_only"], K["merge"])      # bench.py's parse_ms covers both launches
if "scan" in K and "copy" in K: K["stitch"] = add(K["scan"], K["copy"])
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps({k: v for k, v in K.items()}, indent=0)[:600])
