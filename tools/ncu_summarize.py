#!/usr/bin/env python
"""Developer tool: turn the ncu outputs of tools/capture_profiles.sh into the summaries kept under profiles/.
  ncu_summarize.py launches gpurun_out/launches_TAG.csv   -> per-kernel count / mean / share of one device-resident step
  ncu_summarize.py full gpurun_out/raw_TAG.csv [traffic.json] -> table of the --set full capture (+ per-kernel DRAM bytes)"""
import collections
import csv
import json
import re
import sys


def short(name):
    m = re.match(r"(?:void )?(?:tkk::)?([A-Za-z0-9_]+)(<[^>]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    seq = [(short(r[ki]), num(r[vi]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)) for r in rows[1:]]
    # one device-resident step = from a docmark_kernel launch to the next publish_kernel; the bench runs several
    # (warm-up + timed) before the chunked host-buffer leg, whose steps are much shorter: keep the full-size ones
    steps, cur = [], None
    for k, ms in seq:
        if k == "docmark_kernel":
            cur = []
        if cur is not None:
            cur.append((k, ms))
            if k == "publish_kernel":
                steps.append(cur)
                cur = None
    full = [s for s in steps if sum(ms for _, ms in s) > 0.5 * max(sum(ms for _, ms in t) for t in steps)]
    print("# %d launches in the file; %d encode steps, %d of them full-size (device-resident); per-kernel means over those"
          % (len(seq), len(steps), len(full)))
    agg = collections.OrderedDict()
    for s in full:
        for k, ms in s:
            agg.setdefault(k, []).append(ms)
    total = sum(sum(v) for v in agg.values()) / len(full)
    print("kernel | launches per step | ms per step (ncu, serialised) | share of the step")
    for k, v in agg.items():
        print("%s | %d | %.4f | %.1f%%" % (k, len(v) // len(full), sum(v) / len(full), 100 * sum(v) / len(full) / total))
    print("# step total (sum of launches): %.3f ms" % total)
    dec = [ms for k, ms in seq if k.startswith("decode_") or k.startswith("tokmark")]
    if dec:
        print("# decode kernels in the file: %d launches, %.3f ms in total" % (len(dec), sum(dec)))


def full(path, traffic_out=None):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
    units = rows[1]
    print("kernel | " + " | ".join("%s [%s]" % (c, units[idx[c]]) for c in cols if c in idx))
    traffic = {}
    for r in rows[2:]:
        name = short(r[idx["Kernel Name"]])
        print(name + " | " + " | ".join(r[idx[c]] for c in cols if c in idx))
        rd, wr = num(r[idx["dram__bytes_read.sum"]]), num(r[idx["dram__bytes_write.sum"]])
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        b = rd * scale.get(units[idx["dram__bytes_read.sum"]], 1.0) + wr * scale.get(units[idx["dram__bytes_write.sum"]], 1.0)
        key = name.replace("_kernel", "")
        m = re.match(r"lanemerge<(\d+)", key)
        if m:
            key = "lanemerge" + m.group(1)
        traffic[key] = int(b)
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    print("# top stall reasons (warps stalled per issue-active cycle)")
    for r in rows[2:]:
        v = sorted(((num(r[idx[h]]), h.split("stalled_")[1].split("_per")[0]) for h in stalls), reverse=True)[:4]
        print("# %s: %s" % (short(r[idx["Kernel Name"]]), ", ".join("%s %.2f" % (n, x) for x, n in v)))
    print("# DRAM traffic of these kernels: %.2f GB read+write" % (sum(traffic.values()) / 1e9))
    if traffic_out:
        json.dump(traffic, open(traffic_out, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
