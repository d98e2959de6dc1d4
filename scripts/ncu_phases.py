"""Executed instructions of ONE kernel by source-line range ("phase"), out of an ncu report captured with --import-source on:
ncu's SASS page (instruction counts, in program order) joined with the line table of the object file the kernel came from
(`nvdisasm -g`), instruction for instruction; inlined helpers are attributed to the last line of the kernel body seen before them.

    python scripts/ncu_phases.py report.ncu-rep avd_b200/csrc/k1_rows.o k1_rows.cu PIXELS start:name [start:name ...]
e.g. python scripts/ncu_phases.py gpurun_out/r02b_k1rows.ncu-rep avd_b200/csrc/k1_rows.o k1_rows.cu $((256*854*480)) 150:prologue 172:gray 255:laplacian 290:thumbnail 319:cells
"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, obj, src_name, pixels = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
marks = sorted((int(a.split(":")[0]), a.split(":")[1]) for a in sys.argv[5:])
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
sections, cur = [], None                 # one line list per function of the object file
for l in dis.splitlines():
    if l.startswith("//--------------------- .text."):
        sections.append([]); cur = None
        continue
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        cur = int(m.group(2)) if m.group(1).endswith(src_name) else -1
        continue
    if sections and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+", l):
        sections[-1].append(cur)
page = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(page)))
hdr, data = rows[1], rows[2:]
ie = hdr.index("Instructions Executed")
match = [sec for sec in sections if len(sec) == len(data)]     # the function the report's kernel is (same instruction count)
if not match:
    sys.exit(f"no function of the object file has the report's {len(data)} instructions (rebuilt since the capture?): {[len(x) for x in sections]}")
lines = match[0]
acc, last, tot = collections.Counter(), marks[0][1], 0
for i, d in enumerate(data):
    ln = lines[i]
    if ln and ln >= marks[0][0]:
        last = [name for a, name in marks if ln >= a][-1]
    n = int(d[ie]); acc[last] += n; tot += n
for _, name in marks:
    print(f"{name:12s} {100 * acc[name] / tot:5.1f} %  {acc[name] * 32 / pixels:6.2f} lane-instructions / pixel")
print(f"total        {tot} warp instructions = {tot * 32 / pixels:.2f} lane-instructions / pixel")
