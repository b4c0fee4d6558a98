"""Put the per-launch counters of one `ncu --set full` capture into profiles/traffic.json (bench.py reads it).
usage: python tools/update_traffic.py <report.ncu-rep> <workload>/<dtype>/b<batch> [kernel-index] [source-note]"""
import csv, io, json, os, subprocess, sys

rep, key = sys.argv[1], sys.argv[2]
idx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
note = sys.argv[4] if len(sys.argv) > 4 else os.path.basename(rep)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
d = dict(zip(hdr, data[idx]))
u = dict(zip(hdr, units))


def num(name, default=None):
    v = d.get(name)
    if v in (None, ""):
        return default
    x = float(v.replace(",", ""))
    unit = u.get(name, "")
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(unit, 1.0)
    return x * scale


sms = 148
wf = num("l1tex__data_pipe_lsu_wavefronts.sum")
if wf is None:
    wf = num("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed") / 100.0 * num("sm__cycles_elapsed.max") * sms
lts_read = num("lts__t_sectors_op_read.sum")
if lts_read is None:
    lts_read = num("lts__t_sectors_srcunit_tex_op_read.sum", num("lts__t_sectors.sum"))
entry = {
    "dram_bytes": int(num("dram__bytes_read.sum", 0) + num("dram__bytes_write.sum", 0)),
    "lts_read_sectors": int(lts_read),
    "l1_wavefronts": int(wf),
    "l1_hit_rate_pct": num("l1tex__t_sector_hit_rate.pct"),
    "inst_executed": int(num("smsp__inst_executed.sum", 0)),
    "kernel": d.get("Kernel Name"),
    "duration_us_under_ncu": num("gpu__time_duration.sum"),
    "source": note,
}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
try:
    cur = json.load(open(path))
except Exception:
    cur = {}
cur["_source"] = ("per-launch counters of `ncu --set full --clock-control none` captures (rotating cold input sets); written by "
                  "tools/update_traffic.py from the .ncu-rep named in each entry's `source`; summaries of the same captures are "
                  "under profiles/")
cur[key] = entry
json.dump(cur, open(path, "w"), indent=1)
print(key, json.dumps(entry))
