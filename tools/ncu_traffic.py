"""Extract per-launch DRAM traffic of the sampling kernel from an `ncu --set full` report into a small JSON that
bench.py copies into `roofline.traffic`.  Usage: python tools/ncu_traffic.py gpurun_out/prof_sample.ncu-rep profiles/r01_k1_traffic.json"""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
def col(name):
    i = hdr.index(name)
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
    return [float(r[i].replace(",", "")) * scale for r in data]
rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
dur_i = hdr.index("gpu__time_duration.sum")
res = {"kernel": data[0][hdr.index("Kernel Name")], "launches": len(data),
       "dram_bytes_read_per_launch": sum(rd) / len(rd), "dram_bytes_write_per_launch": sum(wr) / len(wr),
       "traffic_bytes_per_launch": (sum(rd) + sum(wr)) / len(rd),
       "ncu_duration_us": [float(r[dur_i].replace(",", "")) for r in data], "ncu_duration_unit": units[dur_i],
       "source": f"ncu --set full --clock-control none, {rep}"}
json.dump(res, open(out, "w"), indent=1)
print(res)
