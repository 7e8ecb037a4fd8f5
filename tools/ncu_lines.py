"""Aggregate an ncu report's warp-stall samples per CUDA source line (needs -lineinfo).
usage: python tools/ncu_lines.py <report.ncu-rep> <kernel-substring> [top]"""
import csv, os, re, subprocess, sys, tempfile
from collections import defaultdict
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "spin_glass_anneal_rl_b200", "libsg_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
txt = None
for f in os.listdir(tmp):
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    if kern in out:
        txt = out.split("\n")
        break
start = [i for i, l in enumerate(txt) if kern in l and l.endswith(":") and not l.startswith(".")][0]
seq, cur, curfile = [], None, None
for l in txt[start + 1:]:
    if l.startswith("//---") and ".text." in l:
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur, curfile = int(m.group(2)), m.group(1).split("/")[-1]
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        seq.append((curfile, cur, m.group(2)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.split("\n")))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr)]
ci = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ci[k]])
    except Exception: return 0.0
assert len(seq) == len(data), (len(seq), len(data))
agg = defaultdict(lambda: [0, 0, defaultdict(float)])
tot = 0
for (fl, ln, sass), r in zip(seq, data):
    s = f(r, "# Samples"); tot += s
    a = agg[(fl, ln)]; a[0] += s; a[1] += f(r, "Instructions Executed")
    for k in hdr:
        if k.startswith("stall_") and "Not Issued" not in k: a[2][k[6:]] += f(r, k)
files = {}
def line(fl, ln):
    for d in [os.path.join(os.path.dirname(so), "csrc"), "/usr/local/cuda/include", "/usr/local/cuda/include/crt"]:
        p = os.path.join(d, fl or "")
        if os.path.exists(p):
            if p not in files: files[p] = open(p, errors="ignore").read().split("\n")
            L = files[p]
            return L[ln - 1].strip()[:60] if ln and ln <= len(L) else ""
    return ""
norm = float(os.environ.get("NORM", "1"))
for (fl, ln), a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    st = sorted(a[2].items(), key=lambda x: -x[1])[:3]
    print(f"{100*a[0]/tot:5.1f}% inst={a[1]/norm:9.1f} {fl}:{ln or 0:<4d} {line(fl, ln):60s} {[(k, int(v)) for k, v in st]}")
print("total instructions /NORM:", sum(a[1] for a in agg.values()) / norm, " samples:", tot)
