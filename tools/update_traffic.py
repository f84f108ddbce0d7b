"""profiles/traffic.json from an `ncu --set full` capture of the bench's roofline kernel:
    python tools/update_traffic.py gpurun_out/<rep>.ncu-rep cfg2 "<how it was captured>"
dram__bytes_read.sum + dram__bytes_write.sum of the captured launch, keyed by the kernel's name (bench.py reads the
entry whose name equals the kernel the library reports, so a changed kernel reads null until re-captured)."""
import csv, io, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, key, how = sys.argv[1], sys.argv[2], sys.argv[3]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, r = rows[0], rows[1], rows[2]
def val(name):
    i = hdr.index(name)
    v, u = float(r[i]), units[i].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
name = r[hdr.index("Kernel Name")]
name = re.sub(r"\(int\)|\(bool\)", "", name).replace("void ", "").replace("frcnn::", "")
name = re.sub(r"\(.*", "", name).replace("<0", "<false").replace(", 0", ", false").replace(", 1>", ", true>")
path = os.path.join(ROOT, "profiles", "traffic.json")
data = json.load(open(path)) if os.path.exists(path) else {}
data[key] = {"kernel": sys.argv[4] if len(sys.argv) > 4 else name, "dram_bytes": int(rd + wr),
             "source": f"{how}: dram__bytes_read.sum {rd / 1e6:.1f} MB + dram__bytes_write.sum {wr / 1e9:.3f} GB, "
                       f"{float(r[hdr.index('gpu__time_duration.sum')]):.1f} us under ncu ({os.path.basename(rep)})"}
json.dump(data, open(path, "w"), indent=1)
print(json.dumps(data[key], indent=1))
