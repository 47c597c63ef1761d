"""Turns the ncu captures gpurun brought back into the committed summaries under profiles/.

    python scripts/ncu_summary.py <tag> <launches.csv> <full.ncu-rep>

Writes profiles/<tag>_launches.md (per-kernel share of the step from the serialised
gpu__time_duration pass), profiles/<tag>_align_kernel.md (the --set full capture of the step's
align_kernel launches: duration, registers, occupancy limits, issue rate, DRAM bytes, stall
reasons) and profiles/traffic.json (dram read+write per step, read by bench.py)."""
import csv
import json
import os
import re
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launches(tag, path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 14 and r[0].isdigit()]
    per = defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = re.sub(r"\(.*", "", r[4]).replace("void ", "")
        name = re.sub(r"<\(int\)(\d+), \(bool\)0>", r"<\1,0>", name)
        per[name][0] += 1
        per[name][1] += float(r[-1].replace(",", "")) / 1e3
    total = sum(v[1] for v in per.values())
    out = [f"# {tag}: launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, bench.py --steps 2 --warmup 3)", "",
           "Per-launch times under ncu are cold-cache and serialised (the real step runs its align_kernel launches",
           "concurrently); compare shares, not absolutes.", "", "| kernel | launches | total µs | share |", "|---|---|---|---|"]
    for name, (n, us) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{name}` | {n} | {us:.1f} | {100 * us / total:.1f} % |")
    fam = sum(v[1] for k, v in per.items() if "align_kernel" in k)
    out += ["", f"align_kernel family: {100 * fam / total:.1f} % of the captured GPU time ({len(rows)} launches captured)."]
    open(os.path.join(ROOT, "profiles", f"{tag}_launches.md"), "w").write("\n".join(out) + "\n")


def full(tag, rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    # the capture window may straddle two steps: keep one launch per (kernel, block, grid) class
    seen = {}
    for r in data:
        seen[(r[idx["Kernel Name"]], r[idx["Block Size"]], r[idx["Grid Size"]])] = r
    data = list(seen.values())
    want = [("gpu__time_duration.sum", "duration"), ("launch__registers_per_thread", "regs"),
            ("launch__occupancy_limit_registers", "occ lim regs (CTAs)"), ("launch__occupancy_limit_shared_mem", "occ lim smem (CTAs)"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
            ("smsp__issue_active.avg.per_cycle_active", "issue/cycle/SMSP"), ("smsp__inst_executed.sum", "warp instr"),
            ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
            ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak")]
    out = [f"# {tag}: `ncu --set full --clock-control none` of one step's launches (align_kernel, backtrack_kernel, finalize_kernel)", "",
           "Each warps-per-segment class is one chain of three launches; under ncu they run one at a time (in the real step",
           "the chains share the GPU), so occupancy and issue rate below are per launch in isolation.", ""]
    out.append("| kernel | block | grid | " + " | ".join(n for _, n in want) + " |")
    out.append("|---|---|---|" + "---|" * len(want))
    rd = wr = 0.0
    inst = 0.0
    per_kernel = defaultdict(lambda: [0.0, 0.0, 0.0, 0.0])      # dram read, dram write, warp instructions, duration us
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    for r in data:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "")
        cells = []
        for m, _ in want:
            v, u = r[idx[m]], units[idx[m]]
            cells.append(f"{float(v.replace(',', '')):.4g} {u}".strip() if v.replace(".", "").replace(",", "").isdigit() else v)
        out.append(f"| `{name}` | {r[idx['Block Size']]} | {r[idx['Grid Size']]} | " + " | ".join(cells) + " |")
        rd += float(r[idx["dram__bytes_read.sum"]].replace(",", "")) * scale.get(units[idx["dram__bytes_read.sum"]], 1.0)
        wr += float(r[idx["dram__bytes_write.sum"]].replace(",", "")) * scale.get(units[idx["dram__bytes_write.sum"]], 1.0)
        inst += float(r[idx["smsp__inst_executed.sum"]].replace(",", ""))
        base = re.sub(r"<.*", "", name).replace("wxb::", "")
        pk = per_kernel[base]
        pk[0] += float(r[idx["dram__bytes_read.sum"]].replace(",", "")) * scale.get(units[idx["dram__bytes_read.sum"]], 1.0)
        pk[1] += float(r[idx["dram__bytes_write.sum"]].replace(",", "")) * scale.get(units[idx["dram__bytes_write.sum"]], 1.0)
        pk[2] += float(r[idx["smsp__inst_executed.sum"]].replace(",", ""))
        pk[3] += float(r[idx["gpu__time_duration.sum"]].replace(",", "")) * {"ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}.get(units[idx["gpu__time_duration.sum"]], 1.0)
    out += ["", "| kernel | dram read MB | dram write MB | warp instructions M | serialised duration us |", "|---|---|---|---|---|"]
    for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1][2]):
        out.append(f"| `{k}` | {v[0] / 1e6:.1f} | {v[1] / 1e6:.1f} | {v[2] / 1e6:.1f} | {v[3]:.1f} |")
    out += ["", f"Step totals over these {len(data)} launches: dram read {rd / 1e6:.1f} MB + write {wr / 1e6:.1f} MB = "
                f"{(rd + wr) / 1e6:.1f} MB; {inst / 1e6:.1f} M warp instructions "
                f"(>= {inst / (148 * 4 * 1.965e9) * 1e6:.0f} µs at 4 issue slots/clk/SM, 1.965 GHz)."]
    # stall reasons of the largest launch
    big = max(data, key=lambda r: float(r[idx["smsp__inst_executed.sum"]].replace(",", "")))
    st = [(h.replace("smsp__pcsamp_warps_issue_stalled_", ""), float(big[i].replace(",", ""))) for h, i in idx.items()
          if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h and big[i].replace(".", "").replace(",", "").isdigit()]
    tot = sum(v for _, v in st) or 1.0
    out += ["", f"Warp-state samples of the largest launch ({big[idx['Block Size']]} x {big[idx['Grid Size']]}): " +
            ", ".join(f"{n} {100 * v / tot:.1f} %" for n, v in sorted(st, key=lambda x: -x[1])[:8]) + "."]
    open(os.path.join(ROOT, "profiles", f"{tag}_align_kernel.md"), "w").write("\n".join(out) + "\n")
    al = per_kernel.get("align_kernel", [0, 0, 0, 0])
    json.dump({"dram_bytes_per_step": al[0] + al[1], "dram_read": al[0], "dram_write": al[1],
               "dram_bytes_per_step_whole_chain": rd + wr, "warp_instructions_per_step": inst,
               "warp_instructions_align_kernel": al[2], "per_kernel": {k: {"dram_read": v[0], "dram_write": v[1], "warp_instructions": v[2]} for k, v in per_kernel.items()},
               "launches": len(data), "source": f"profiles/{tag}_align_kernel.md (ncu --set full, {os.path.basename(rep)})"},
              open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    tag, lcsv, rep = sys.argv[1:4]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    launches(tag, lcsv)
    full(tag, rep)
    print("wrote profiles/", tag)
