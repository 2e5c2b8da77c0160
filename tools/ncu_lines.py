"""Per-source-line aggregation of an ncu report: joins `ncu -i rep --page source --csv` (SASS rows with sample / instruction
counts) with `nvdisasm --print-line-info` of the cubin the report was taken from.
usage: ncu_lines.py <report.ncu-rep> <object-or-cubin> <kernel-substring> [top=40]"""
import csv, os, re, subprocess, sys, tempfile

rep, obj, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
if not obj.endswith(".cubin"):
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    obj = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", obj], capture_output=True, text=True).stdout.splitlines()
line_of, cur, inside = {}, None, False
for ln in dis:
    if ln.startswith("//---") and ".text." in ln:
        inside = kname in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[h], rows[h + 1:]
ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
base = int(data[0][ia], 16)
agg, tot_s, tot_e = {}, 0, 0
for r in data:
    if len(r) <= iex:
        continue
    off = int(r[ia], 16) - base
    key = line_of.get(off, ("?", 0))
    s, e = int(r[isamp]), int(r[iex])
    a = agg.setdefault(key, [0, 0])
    a[0] += s; a[1] += e
    tot_s += s; tot_e += e
src = {}
print(f"total samples {tot_s}, warp instructions {tot_e}")
for key, (s, e) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    f, l = key
    text = ""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "slam_toolbox_b200", "csrc", f)
    if os.path.exists(path):
        if f not in src:
            src[f] = open(path).read().splitlines()
        if 0 < l <= len(src[f]):
            text = src[f][l - 1].strip()[:110]
    print(f"{100 * s / max(tot_s, 1):5.1f}% samp {100 * e / max(tot_e, 1):5.1f}% inst  {f}:{l}  {text}")
