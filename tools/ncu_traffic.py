"""Per-kernel-class DRAM traffic / duration / tensor-pipe summary of one eval from `ncu --page raw --csv` output.
usage: ncu_traffic.py raw.csv out.json"""
import csv, json, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
while rows and (len(rows[0]) < 5 or rows[0][0] != "ID"):      # skip ncu's ==PROF== log lines (--log-file mode)
    rows.pop(0)
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def val(r, k, scale_units=True):
    if k not in ix or r[ix[k]] in ("", "n/a"): return 0.0
    v = float(r[ix[k]].replace(",", ""))
    u = units[ix[k]]
    mul = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0, "second": 1.0,
           "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9}.get(u, 1.0)
    return v * mul
def klass(name):
    name = name.replace("void ", "").replace("rho::", "")
    m = re.match(r"([A-Za-z0-9_]+)", name)
    return m.group(1) if m else name
agg = collections.OrderedDict()
for r in data:
    k = klass(r[ix["Kernel Name"]])
    a = agg.setdefault(k, {"launches": 0, "time_s": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "tensor_pct_time_weighted": 0.0})
    t = val(r, "gpu__time_duration.sum")
    a["launches"] += 1
    a["time_s"] += t
    a["dram_read_bytes"] += val(r, "dram__bytes_read.sum")
    a["dram_write_bytes"] += val(r, "dram__bytes_write.sum")
    a["tensor_pct_time_weighted"] += t * val(r, "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active")
for a in agg.values():
    a["tensor_pipe_active_pct"] = a.pop("tensor_pct_time_weighted") / a["time_s"] if a["time_s"] else 0.0
    a["dram_gbs"] = (a["dram_read_bytes"] + a["dram_write_bytes"]) / a["time_s"] / 1e9 if a["time_s"] else 0.0
conv = [v for k, v in agg.items() if k.startswith("conv_tc_kernel") or k.startswith("conv_halo_kernel")]
out = {"per_kernel": agg,
       "conv_tcgen05": {"launches": sum(c["launches"] for c in conv), "time_s": sum(c["time_s"] for c in conv),
                        "dram_bytes": sum(c["dram_read_bytes"] + c["dram_write_bytes"] for c in conv)}}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out["conv_tcgen05"]))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["time_s"]):
    print(f"{a['time_s']*1e3:8.3f} ms {a['launches']:4d}  tensor {a['tensor_pipe_active_pct']:5.1f}%  dram {a['dram_gbs']:7.1f} GB/s  {k}")
