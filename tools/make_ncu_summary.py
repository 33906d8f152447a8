"""profiles/ncu_summary.json from an `ncu --page raw --csv` export: per kernel the DRAM bytes per launch (bench.py's
roofline.traffic), duration and pipe utilisation.  python tools/make_ncu_summary.py raw.csv "source description" f64 > profiles/ncu_summary.json"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
def val(d, key):
    v = float(d[idx[key]].replace(",", ""))
    u = units[idx[key]]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}.get(u, 1.0)
out = {"source": sys.argv[2], "dtype": sys.argv[3] if len(sys.argv) > 3 else "f64"}
for d in data:
    name = d[idx["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
    e = {"dram_bytes_per_launch": val(d, "dram__bytes_read.sum") + val(d, "dram__bytes_write.sum"),
         "duration_us": val(d, "gpu__time_duration.sum"), "grid": d[idx["launch__grid_size"]], "block": d[idx["launch__block_size"]],
         "registers": d[idx["launch__registers_per_thread"]]}
    for k, key in (("tensor_fp64_pct", "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed"),
                   ("fp64_pipe_pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                   ("tensor_pipe_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                   ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                   ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                   ("l2_hit_pct", "lts__t_sector_hit_rate.pct")):
        if key in idx:
            try:
                e[k] = float(d[idx[key]])
            except ValueError:
                pass
    out[name] = e
print(json.dumps(out, indent=1))
