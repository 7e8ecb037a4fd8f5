"""Turn the round's ncu outputs (gpurun_out/) into the committed summaries under profiles/.
usage: python tools/summarize_profile.py <round-tag> <launches.csv> <prof.ncu-rep> <kernel-substring> <bench.log> [probe.ncu-rep]
(<kernel-substring>: the mangled template instance for the per-line stall table, e.g.
sweep_tc_kernelILi3ELb0ELi4E)"""
import collections, csv, json, os, re, subprocess, sys
tag, launches, rep, kern, benchlog = sys.argv[1:6]
probe = sys.argv[6] if len(sys.argv) > 6 else None
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(root, "profiles")
os.makedirs(out_dir, exist_ok=True)

# ---- launch list
rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "s": 1e3}
agg = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki]).strip()
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", "")) * scale[r[ui]]
tot = sum(a[1] for a in agg.values())
with open(os.path.join(out_dir, f"{tag}_bench_launches.txt"), "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 5 --warmup 3 --ttt-budget 0 --cpu-budget 2\n")
    f.write(f"# every kernel launched by the command (cold-cache, serialised: compare shares)\n")
    f.write(f"# total device time {tot:.3f} ms over {sum(a[0] for a in agg.values())} launches\n")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f"{a[0]:5d} launches {a[1]:10.3f} ms {100 * a[1] / tot:6.2f}%  {a[1] / a[0]:9.4f} ms/launch  {k}\n")

# ---- full capture of the dominant kernel
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, u, d = rr[0], rr[1], rr[2]
def get(name):
    i = h.index(name)
    return d[i], u[i]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "lts__t_sectors.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max"]
def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v.replace(",", "")) * m[unit]
with open(os.path.join(out_dir, f"{tag}_sweep_tc_ncu.txt"), "w") as f:
    f.write(f"# ncu --set full --clock-control none --import-source on -k regex:sweep_tc  python tools/prof.py tc\n")
    f.write(f"# (SK N=4096, 8192 replicas, 10 sweeps per launch, 3 planes: the bench launch shape)\n")
    f.write(f"kernel: {d[h.index('Kernel Name')]}\n")
    for w in want:
        if w in h:
            v, un = get(w)
            f.write(f"{w:70s} {v:>18s} {un}\n")
    f.write("\n# warp-stall samples per source line (tools/ncu_lines.py)\n")
    lines = subprocess.run([sys.executable, os.path.join(root, "tools", "ncu_lines.py"), rep, kern, "24"],
                           capture_output=True, text=True).stdout
    f.write(lines)
rd, wr = to_bytes(*get("dram__bytes_read.sum")), to_bytes(*get("dram__bytes_write.sum"))
dur_v, dur_u = get("gpu__time_duration.sum")
lts = None
for cand in ("lts__t_bytes.sum",):
    if cand in h:
        lts = to_bytes(*get(cand))
if lts is None and "lts__t_sectors.sum" in h:
    lts = float(get("lts__t_sectors.sum")[0].replace(",", "")) * 32.0
json.dump({"kernel": "sg::sweep_tc_kernel", "source": f"profiles/{tag}_sweep_tc_ncu.txt (ncu --set full of tools/prof.py tc)",
           "shape": "SK N=4096, 8192 replicas, 10 sweeps per launch, 3 bf16 planes",
           "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
           "lts_bytes_per_launch": lts,
           "ncu_duration": f"{dur_v} {dur_u}"},
          open(os.path.join(out_dir, f"{tag}_sweep_tc_traffic.json"), "w"), indent=1)

# ---- the L2 stream probe (the roofline denominator)
if probe and os.path.exists(probe):
    praw = subprocess.run(["ncu", "-i", probe, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    pr = list(csv.reader(praw.splitlines()))
    ph, pu = pr[0], pr[1]
    with open(os.path.join(out_dir, f"{tag}_tma_probe_ncu.txt"), "w") as f:
        f.write("# ncu --set full --clock-control none -k regex:tma_probe  python tools/prof.py tma\n")
        f.write("# sg_measure_tma_stream: every SM pulls the same L2-resident 65 MB buffer in the same order through a\n")
        f.write("# shared-memory ring with TMA bulk copies, no compute (launch 1: 17.9 KB copies x 8 stages, 2: 48 KB x 4)\n")
        for row in pr[2:]:
            f.write(f"kernel: {row[ph.index('Kernel Name')]}\n")
            for w in ["gpu__time_duration.sum", "lts__t_bytes.sum", "lts__t_sectors.sum", "dram__bytes_read.sum",
                      "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
                      "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__grid_size", "launch__block_size"]:
                if w in ph:
                    i = ph.index(w)
                    f.write(f"  {w:66s} {row[i]:>18s} {pu[i]}\n")
    print(open(os.path.join(out_dir, f"{tag}_tma_probe_ncu.txt")).read())

# ---- the bench line of the same command
line = [l for l in open(benchlog) if l.startswith("{")][-1]
open(os.path.join(out_dir, f"{tag}_bench.json"), "w").write(line)
print(open(os.path.join(out_dir, f"{tag}_bench_launches.txt")).read())
print(open(os.path.join(out_dir, f"{tag}_sweep_tc_ncu.txt")).read()[:2500])
