"""dram traffic per launch of one kernel from an `ncu --set full` report -> JSON record read by bench.py (roofline.traffic):
    python tools/ncu_traffic.py REPORT.ncu-rep KERNEL_SUBSTR BATCH OUT.json"""
import csv, json, subprocess, sys
rep, kern, batch, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
sel = [r for r in rows[2:] if kern in r[ix["Kernel Name"]]]
def col(name):
    return [float(r[ix[name]].replace(",", "")) * scale[units[ix[name]]] for r in sel]
rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
ms = [float(r[ix["gpu__time_duration.sum"]].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[units[ix["gpu__time_duration.sum"]]] for r in sel]
rec = {"kernel": sel[0][ix["Kernel Name"]].split("(")[0], "batch": batch, "launches_captured": len(sel),
       "dram_bytes_read": sum(rd) / len(rd), "dram_bytes_write": sum(wr) / len(wr), "launch_ms_under_ncu": sum(ms) / len(ms),
       "source": "ncu --set full --clock-control none of `python tools/profile_step.py %d 2` (%s)" % (batch, rep.split("/")[-1])}
json.dump(rec, open(out, "w"), indent=1)
print(rec)
