#!/usr/bin/env python
"""Turn the raw ncu CSV exports that tools/profile_gpu.sh leaves in gpurun_out/ into the small, tracked evidence files
under profiles/ (what DESIGN.md and bench.py's `roofline.traffic` cite):

    python tools/summarise_profiles.py r01

    profiles/<tag>_uvd_launches.csv, <tag>_kron_launches.csv    launch lists (kernel, grid, block, device time)
    profiles/<tag>_ncu_full_summary.csv                         one row per --set full capture: time, DRAM bytes,
                                                                DRAM/SM/tensor-pipe/LSU utilisation, registers, issue
    profiles/<tag>_hotspots.txt                                 top stall instructions of the dominant kernels
    profiles/<tag>_traffic.json                                 per-kernel DRAM traffic per launch (read by bench.py)
"""
from __future__ import annotations

import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out")
DST = os.path.join(ROOT, "profiles")


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("psgd::uvd::", "").replace("psgd::tc::", "").replace("psgd::kron::", "").replace("psgd::", "")


def launches(tag: str, which: str):
    path = os.path.join(SRC, f"{tag}_{which}_launches.csv")
    if not os.path.exists(path):
        return None
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr, data = rows[0], rows[1:]
    ix = {h: i for i, h in enumerate(hdr)}
    out = os.path.join(DST, f"{tag}_{which}_launches.csv")
    tot = {}
    with open(out, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["# ncu --metrics gpu__time_duration.sum --clock-control none (tools/profile_gpu.sh): every launch of this "
                    "repo's kernels in the timed region; cold-cache, serialised -- compare SHARES, not absolutes"])
        w.writerow(["launch", "kernel", "grid", "block", "duration_us"])
        for r in data:
            k = short(r[ix["Kernel Name"]])
            us = float(r[ix["Metric Value"]]) / 1e3
            w.writerow([r[ix["ID"]], k, r[ix["Grid Size"]], r[ix["Block Size"]], f"{us:.2f}"])
            t = tot.setdefault(k, [0, 0.0])
            t[0] += 1
            t[1] += us
        w.writerow([])
        w.writerow(["# totals", "kernel", "launches", "total_us", "share"])
        total = sum(v[1] for v in tot.values()) or 1.0
        for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            w.writerow(["", k, n, f"{us:.1f}", f"{us / total:.4f}"])
    return tot


RAW_COLS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_pct"),
    ("sm__inst_executed_pipe_tensor_op_utcmma.sum", "utcmma_inst"),   # not in every ncu version
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lsu_shared_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp_inst"),
]


def to_bytes(val: str, unit: str) -> float:
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1.0)
    return float(val) * mult


def to_us(val: str, unit: str) -> float:
    mult = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(unit, 1.0)
    return float(val) * mult


def full_rows(tag: str, name: str):
    path = os.path.join(SRC, f"{tag}_{name}_raw.csv")
    if not os.path.exists(path):
        return []
    rows = list(csv.reader(open(path, errors="replace")))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in data:
        d = {"capture": name, "kernel": short(r[ix["Kernel Name"]])}
        for col, key in RAW_COLS:
            # the TriageCompute-prefixed duplicates carry a section prefix; take the plain metric name
            i = ix.get(col)
            if i is None or r[i] == "":
                d[key] = ""
                continue
            if key == "time":
                d["time_us"] = round(to_us(r[i], units[i]), 2)
            elif key in ("dram_read", "dram_write"):
                d[key + "_MB"] = round(to_bytes(r[i], units[i]) / 1e6, 2)
            else:
                d[key] = r[i]
        out.append(d)
    return out


def hotspots(tag: str, name: str, top: int = 25):
    path = os.path.join(SRC, f"{tag}_{name}_source.csv")
    if not os.path.exists(path):
        return ""
    rows = list(csv.reader(open(path, errors="replace")))
    if not any(r and r[0] == "Address" for r in rows):
        return ""
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    kernel = rows[0][1] if len(rows[0]) > 1 else name
    cols = rows[hdr]
    ix = {h: i for i, h in enumerate(cols)}
    data = [r for r in rows[hdr + 1:] if len(r) > 5 and r[2].isdigit()]
    total = sum(int(r[2]) for r in data) or 1
    stall_cols = [h for h in cols if h.startswith("stall_") and "Not Issued" not in h]
    lines = [f"== {name}: {short(kernel)} -- warp-state samples per SASS instruction ({total} samples; top {top})",
             f"{'samples':>8} {'share':>6} {'executed':>10}  instruction / dominant stall reasons"]
    for r in sorted(data, key=lambda r: -int(r[2]))[:top]:
        st = sorted(((h[6:], int(r[ix[h]])) for h in stall_cols if r[ix[h]].isdigit() and int(r[ix[h]]) > 0), key=lambda kv: -kv[1])[:3]
        lines.append(f"{r[2]:>8} {int(r[2]) / total:6.1%} {r[5]:>10}  {r[1].strip()[:80]:<80} {st}")
    return "\n".join(lines) + "\n"


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(DST, exist_ok=True)
    for which in ("uvd", "kron"):
        launches(tag, which)
    allrows = []
    for name in ("uvd_full", "kron_full_head", "kron_full_tail", "kron_full", "gemm4096", "splu_full", "ns_full"):
        allrows += full_rows(tag, name)
    if allrows:
        keys = ["capture", "kernel", "time_us", "dram_read_MB", "dram_write_MB", "dram_pct", "sm_pct", "tensor_pipe_pct",
                "lsu_shared_pct", "issue_active_pct", "regs", "grid", "block", "warp_inst"]
        with open(os.path.join(DST, f"{tag}_ncu_full_summary.csv"), "w", newline="") as fh:
            fh.write("# ncu --set full --clock-control none --import-source on (tools/profile_gpu.sh); one row per captured launch.\n"
                     "# uvd_full: the sweeps of one UVd update+apply at N=1e8, r=10 (r01: five sweeps of the two-call form; later tags: the three sweeps of the fused call).  kron_full_head/tail: launches 0-1 and the\n"
                     "# last 8 GEMM launches of one Kron step (6 layers per grouped launch); kron_full (r02 on): ALL tcgen05 GEMM launches of one step, in launch order.  gemm4096: one dense 4096^3 product.\n")
            w = csv.DictWriter(fh, fieldnames=keys, extrasaction="ignore")
            w.writeheader()
            for d in allrows:
                w.writerow(d)
        # DRAM traffic per launch for bench.py's roofline.traffic; entries of earlier captures (other kernels) are kept
        traffic = {}
        prev = sorted(f for f in os.listdir(DST) if f.endswith("_traffic.json") and not f.startswith(tag + "_"))
        if prev:
            traffic.update(json.load(open(os.path.join(DST, prev[-1]))).get("bytes_per_launch", {}))
        names = {"gram_sweep_kernel<10, 0>": "uvd_gram_update", "map_sweep_kernel<10, 0>": "uvd_map_update2",
                 "map_sweep_kernel<10, 1>": "uvd_map_update3", "map_sweep_kernel<10, 2>": "uvd_map_update3",
                 "gram_sweep_kernel<10, 1>": "uvd_gram_apply", "map_sweep_kernel<10, 3>": "uvd_map_apply",
                 "map_sweep_kernel<10, 7>": "uvd_map_fused", "map_sweep_kernel<10, 8>": "uvd_map_fused",
                 "map_sweep_kernel<10, 9>": "uvd_map_updapp", "map_sweep_kernel<10, 10>": "uvd_map_updapp",
                 "map_sweep_kernel<10, 11>": "uvd_map_apply_d", "d_update_kernel": "uvd_d_update"}
        for d in allrows:
            if d["capture"] == "uvd_full" and d["kernel"] in names and d.get("dram_read_MB") != "":
                traffic[names[d["kernel"]]] = int((d["dram_read_MB"] + d["dram_write_MB"]) * 1e6)
        kron = [d for d in allrows if d["capture"].startswith("kron_full") and d.get("dram_read_MB") != ""]
        if kron:
            traffic["kron_gemm_tc_per_launch_6_layers_mean"] = int(sum(d["dram_read_MB"] + d["dram_write_MB"] for d in kron) / len(kron) * 1e6)
        g = [d for d in allrows if d["capture"] == "gemm4096" and d.get("dram_read_MB") != ""]
        if g:
            traffic["gemm_tc_dense_4096"] = int((g[0]["dram_read_MB"] + g[0]["dram_write_MB"]) * 1e6)
        json.dump({"source": f"profiles/{tag}_ncu_full_summary.csv (dram__bytes_read.sum + dram__bytes_write.sum per launch; "
                             "kernels not re-captured under this tag keep the previous capture's figure)",
                   "bytes_per_launch": traffic}, open(os.path.join(DST, f"{tag}_traffic.json"), "w"), indent=1)
    hs = "".join(hotspots(tag, n) for n in ("uvd_full", "uvd_map", "kron_full_tail", "kron_full", "gemm4096", "gemm4096_ts", "splu_full", "ns_full"))
    if hs:
        open(os.path.join(DST, f"{tag}_hotspots.txt"), "w").write(
            "# ncu --page source (per-instruction warp-state sampling) of the dominant kernels; -lineinfo builds.\n" + hs)
    print("profiles/:", sorted(os.listdir(DST)))


if __name__ == "__main__":
    main()
