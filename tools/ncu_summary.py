"""Summarise `ncu -i X.ncu-rep --page raw --csv` output: per launch duration, DRAM bytes, tensor-pipe and DRAM utilisation."""
import csv, sys, json
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
if len(sys.argv) > 2 and sys.argv[2] == "--grep":
    for h in hdr:
        if sys.argv[3] in h: print(h, units[ix[h]], [r[ix[h]] for r in data][:4])
    sys.exit(0)
out = []
for r in data:
    d = {"kernel": r[ix["Kernel Name"]][:110], "grid": r[ix["Grid Size"]], "block": r[ix["Block Size"]]}
    for k in KEYS:
        if k in ix: d[k] = r[ix[k]] + " " + units[ix[k]]
    out.append(d)
print(json.dumps(out, indent=1))
