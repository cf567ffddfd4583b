"""Turn ncu outputs into the small text summaries committed under profiles/.
  python tools/summarize_ncu.py launches gpurun_out/r01_launches.csv  > profiles/r01_launches_summary.md
  python tools/summarize_ncu.py full     gpurun_out/r01_gemm2_ffn.ncu-rep > profiles/r01_gemm2_ffn_summary.md
"""
import collections
import csv
import io
import subprocess
import sys


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    tot = collections.defaultdict(lambda: [0.0, 0])
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1}.get(unit, 1)
        name = r["Kernel Name"].split("(")[0][:90]
        tot[name][0] += ns
        tot[name][1] += 1
    total = sum(v[0] for v in tot.values())
    n = sum(v[1] for v in tot.values())
    print(f"# ncu launch list: {n} launches, {total/1e6:.2f} ms summed device time (cold-cache, serialised: compare SHARES)\n")
    mine = sum(v[0] for k, v in tot.items() if "ofk::" in k or k.startswith("gemm") or "attn_" in k or "ln_" in k)
    print(f"libofk kernels: {mine/1e6:.2f} ms = {100*mine/total:.1f} % of the captured launches\n")
    print("| share | ms | launches | kernel |\n|---|---|---|---|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0])[:40]:
        print(f"| {100*v[0]/total:5.1f} % | {v[0]/1e6:8.3f} | {v[1]} | `{k}` |")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
            "launch__block_size", "launch__cluster_size"]
    want = [w for w in want if w in idx]
    print(f"# ncu --set full: {path}\n")
    print("| kernel | " + " | ".join(f"{w} [{units[idx[w]]}]" for w in want) + " |")
    print("|---|" + "---|" * len(want))
    for r in rows[2:]:
        print("| `" + r[idx["Kernel Name"]][:48] + "` | " + " | ".join(r[idx[w]] for w in want) + " |")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
