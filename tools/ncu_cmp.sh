python tools/dump_csr.py 10000 1 /tmp/c1 > /dev/null
export ELECTOR_PIPELINE_CHUNKS=1 ELECTOR_PIPELINE_WORKERS=1
for v in new old; do
  if [ $v = old ]; then mkdir -p /tmp/oldlib; cp elector_b200/libelector_poa_old.so /tmp/oldlib/libelector_poa.so; export LD_LIBRARY_PATH=/tmp/oldlib; fi
  timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 120 --csv --log-file /tmp/l_$v.csv elector_b200/bin/pipe_driver /tmp/c1 2 > /dev/null 2>&1
  python - /tmp/l_$v.csv $v <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]; ik, im, iv, ig = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Grid Size")
d = {}
order = []
for r in rows[1:]:
    key = (r[h.index("ID")],)
    d.setdefault(key, {"k": r[ik].split("(")[0].replace("void ", "").replace("elector::", "")[:44], "g": r[ig]})[r[im]] = float(r[iv].replace(",", ""))
ks = list(d.values())
half = len(ks) // 2
tot = {}
for e in ks[half:]:   # second call
    if "poa_dp" in e["k"]: print("%-5s %-46s %-14s %9.1f us %12.0f inst" % (sys.argv[2], e["k"], e["g"], e.get("gpu__time_duration.sum", 0) / 1e3, e.get("smsp__inst_executed.sum", 0)))
    tot[e["k"]] = tot.get(e["k"], 0) + e.get("gpu__time_duration.sum", 0) / 1e3
print(sys.argv[2], "total us by kernel:", {k: round(v) for k, v in tot.items()})
PY
done
