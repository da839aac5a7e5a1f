"""Turn gpurun_out/ ncu artefacts into the small tracked summaries under profiles/.
usage: python tools/summarize_profiles.py <round-tag> <launches.csv> [name=report.ncu-rep ...]"""
import csv
import collections
import re
import subprocess
import sys

tag, launches = sys.argv[1], sys.argv[2]
reports = dict(a.split("=", 1) for a in sys.argv[3:])


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    return name[:90]


rows = [r for r in csv.reader(l for l in open(launches) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows:
    k = short(r[ki])
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", ""))
tot = sum(v[1] for v in agg.values())
with open(f"profiles/{tag}_launches_summary.md", "w") as f:
    f.write(f"# {tag}: ncu launch list of `python bench.py --steps 2 --warmup 1` (gpu__time_duration.sum, --clock-control none)\n\n")
    f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n")
    f.write(f"{len(rows)} launches captured, {tot / 1e6:.2f} ms total.\n\n| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {c} | {t / 1e3:.1f} | {t / tot:.3f} |\n")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__shared_mem_per_block_dynamic"]
for name, rep in reports.items():
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(out.splitlines()))
    h, units, vals = rr[0], rr[1], rr[2]
    with open(f"profiles/{tag}_{name}_ncu.md", "w") as f:
        f.write(f"# {tag}: `ncu --set full --clock-control none` of {name} ({rr[2][h.index('Kernel Name')][:80]})\n\n| metric | unit | value |\n|---|---|---:|\n")
        for a, u, v in zip(h, units, vals):
            if a in WANT:
                f.write(f"| {a} | {u} | {v} |\n")
print("ok")
