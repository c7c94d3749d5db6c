#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries kept under profiles/.

    python scripts/summarize_ncu.py launches gpurun_out/launches_r1.csv  > profiles/launches_r1.md
    python scripts/summarize_ncu.py kernel   gpurun_out/prof.ncu-rep     > profiles/scan_kernel_r1.md
The `kernel` mode also writes profiles/scan_kernel_traffic.json (dram bytes per launch), which
bench.py reports as roofline.traffic."""
import csv
import json
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def short(name: str) -> str:
    if "at::" in name or "vectorized_elementwise" in name:
        return "torch: " + ("randn" if "normal" in name else "copy / fill / elementwise")
    name = name.replace("void ", "")
    if name.startswith("cutlass") or "gemm" in name.lower():
        return "cuBLAS TF32 GEMM 8192^3 (bench.py's live yardstick, outside the timed region)"
    if "flat_scan_tc_kernel<" in name:     # the template arguments tell the seeding pre-pass (..., 32, 0, 16) from the main scan
        args = name.split("flat_scan_tc_kernel<")[1].split(">")[0].replace("(int)", "").replace("(bool)", "").replace(" ", "")
        parts = args.split(",")
        seed = len(parts) >= 5 and parts[4] != "0"      # kept minima per item (16 / 32; "1" in captures before round 2's last build)
        ham = parts[5] if len(parts) >= 6 else "0"
        kind = " (seeding pre-pass)" if seed else {"1": " (Hamming scan: bf16 codes, collect within the bound)"}.get(ham, " (main scan)")
        return "flat_scan_tc_kernel<" + args + ">" + kind
    for cut in ("<", "("):
        if cut in name:
            name = name.split(cut)[0]
    return name.split("::")[-1]


def launches(path: str) -> None:
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = OrderedDict()
    for r in rows[start + 1:]:
        if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        unit = r[ix["Metric Unit"]]
        val = float(r[ix["Metric Value"]].replace(",", ""))
        us = val / 1e3 if unit in ("ns", "nsecond") else val if unit in ("us", "usecond") else val * 1e3
        name = short(r[ix["Kernel Name"]])
        if name.endswith("(main scan)") and us < 100.0:      # same instance, launched again for the redo pass
            name = name.replace("(main scan)", "(redo launch: no query tile flagged)")
        agg.setdefault(name, []).append(us)
    total = sum(sum(v) for v in agg.values())
    print(f"# Launch list ({os.path.basename(path)})\n")
    print("`ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py --steps 3 --warmup 3 "
          "--no-cpu-baseline`; per-launch times are cold-cache and serialised - compare shares, not absolutes.\n")
    ours = {k: v for k, v in agg.items() if not (k.startswith("torch:") or k.startswith("cuBLAS"))}
    ours_total = sum(sum(v) for v in ours.values()) or 1.0
    print("| kernel | launches | mean us | total us | share of all launches | share of the search step |\n|---|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        step = f"{100 * sum(v) / ours_total:.1f} %" if k in ours else "- (outside the step)"
        print(f"| `{k}` | {len(v)} | {sum(v) / len(v):,.1f} | {sum(v):,.1f} | {100 * sum(v) / total:.1f} % | {step} |")


KEYS = [
    "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpc__cycles_elapsed.avg.per_second",
]


def kernel(path: str) -> None:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# `ncu --set full --clock-control none` ({os.path.basename(path)})\n")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"## `{short(d['Kernel Name'])}` grid {d.get('Grid Size')} block {d.get('Block Size')}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for h, u in zip(hdr, units):
            if any(h.endswith(k) or h == k for k in KEYS) or "pipe_tensor" in h and "pct" in h:
                print(f"| `{h}` | {d[h]} | {u} |")
        rd = [float(d[h].replace(",", "")) * (1e9 if units[hdr.index(h)] == "Gbyte" else 1e6 if units[hdr.index(h)] == "Mbyte" else 1)
              for h in hdr if h == "dram__bytes_read.sum"]
        wr = [float(d[h].replace(",", "")) * (1e9 if units[hdr.index(h)] == "Gbyte" else 1e6 if units[hdr.index(h)] == "Mbyte" else 1)
              for h in hdr if h == "dram__bytes_write.sum"]
        if rd and wr:
            secs = [float(d[h].replace(",", "")) * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(units[hdr.index(h)].replace("second", "s").replace("mss", "ms").replace("uss", "us").replace("nss", "ns"), 1e-3)
                    for h in hdr if h == "gpu__time_duration.sum"]
            if secs and secs[0] > 0:
                print(f"| DRAM read + write per launch / duration | {(rd[0] + wr[0]) / secs[0] / 1e9:.1f} | GB/s |")
        if rd and wr and "flat_scan_tc" in d["Kernel Name"]:
            with open(os.path.join(ROOT, "profiles", "scan_kernel_traffic.json"), "w") as f:
                json.dump({"kernel": short(d["Kernel Name"]), "dram_bytes_per_launch": rd[0] + wr[0], "source": os.path.basename(path)}, f)
        print()


def source(path: str) -> None:
    """Warp-state sample shares and the hottest SASS instructions of every kernel in the report."""
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    i = 0
    while i < len(rows):
        if not rows[i] or rows[i][0] != "Kernel Name":
            i += 1
            continue
        kname, hdr = rows[i][1], rows[i + 1]
        j = i + 2
        while j < len(rows) and rows[j] and rows[j][0] != "Kernel Name":
            j += 1
        data = [r for r in rows[i + 2:j] if len(r) == len(hdr)]
        ix = {h: n for n, h in enumerate(hdr)}
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        tot = sum(int(r[ix["# Samples"]] or 0) for r in data) or 1
        print(f"## Warp-state samples, `{short(kname)}` (source page, all warps, {tot} samples)\n")
        print("| stall reason | share |\n|---|---|")
        share = {h: sum(int(r[ix[h]] or 0) for r in data) for h in stalls}
        st = sum(share.values()) or 1
        for h, v in sorted(share.items(), key=lambda kv: -kv[1]):
            if v * 100.0 / st >= 1.0:
                print(f"| {h[6:]} | {v * 100.0 / st:.1f} % |")
        print("\n### Hottest instructions\n\n| samples | executed | SASS | top stall |\n|---|---|---|---|")
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:14]:
            top = max(stalls, key=lambda h: int(r[ix[h]] or 0))
            print(f"| {int(r[ix['# Samples']] or 0) * 100.0 / tot:.1f} % | {r[ix['Instructions Executed']]} | `{r[ix['Source']][:70]}` | {top[6:]} |")
        print()
        i = j


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel, "source": source}[sys.argv[1]](sys.argv[2])
