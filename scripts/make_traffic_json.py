"""Build profiles/traffic.json from an ncu metrics CSV of scripts/prof_round2.py:

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:"siren_rev_kernel|isect_|contour_" -c 80 --csv --log-file gpurun_out/traffic.csv python scripts/prof_round2.py

bench.py reports `roofline.traffic` from this file only while the kernel source is byte-identical to the profiled one
(sha256 of the .cu file is stored next to the number)."""
import csv, hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "traffic.csv")
rows = list(csv.reader(open(src)))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[h]
ki, mi, vi, ui = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("Metric Unit")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
launches = {}
for r in rows[h + 1:]:
    if len(r) <= vi or "bytes" not in r[mi]:
        continue
    launches.setdefault(int(r[0]), [r[ki], 0.0])[1] += float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
order = [launches[k] for k in sorted(launches)]
def last_call(prefix, count_per_call):
    sel = [b for n, b in order if prefix in n]
    return sum(sel[-count_per_call:]) if len(sel) >= count_per_call else None
sha = lambda f: hashlib.sha256(open(os.path.join(ROOT, "inf-3dp_b200", "csrc", f), "rb").read()).hexdigest()
out = {
    "siren_rev": {"dram_bytes": last_call("siren_rev_kernel", 1), "source_sha256": sha("siren_rev.cu"),
                  "what": "one launch of siren_rev_kernel<2,0> on the 512^3 grid (134 217 728 points), rows output"},
    "intersection": {"dram_bytes": last_call("isect_", 4), "source_sha256": sha("intersection.cu"),
                     "what": "prep + init + elect + cell at D = 255, Ns = 500, N1 = N2 = 128"},
    "contours_dense": {"dram_bytes": last_call("contour_", 9), "source_sha256": sha("contours.cu"),
                       "what": "the nine launches of one dense extract_isocontours call, F = 1 M, K = 200"},
    "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none (scripts/make_traffic_json.py)",
}
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
