"""Turns the ncu captures brought back in gpurun_out/ into small tracked summaries under profiles/:
  profiles/<tag>_launches.md   kernel shares from the gpu__time_duration launch list
  profiles/<tag>_kernels.md    per-kernel metrics from the --set full capture (dram bytes, throughput %, occupancy ...)
  profiles/traffic.json        dram bytes per launch of each kernel (read by bench.py -> roofline.traffic)
Usage: python tools/summarize_profiles.py r01 gpurun_out/r01_launches.csv gpurun_out/r01_env_advance.ncu-rep gpurun_out/r01_replay.ncu-rep
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_uniform.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "Tbyte": 1e12}


def short(name):
    name = name.replace("void ", "").replace("qlc::", "")
    return name.split("(")[0]


def launches(tag, path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            agg.setdefault(short(r[ki]), []).append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    tot = sum(sum(v) for v in agg.values())
    out = ["# %s — launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, first %d launches of `%s`)" % (tag, sum(len(v) for v in agg.values()), os.environ.get("PROFILE_CMD", "python bench.py --steps 5 --warmup 3 --no-cpu-baseline")),
           "", "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.", "",
           "| kernel | launches | total us | avg us | share |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append("| `%s` | %d | %.1f | %.2f | %.1f%% |" % (k, len(v), sum(v) / 1e3, sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
    open(os.path.join(ROOT, "profiles", tag + "_launches.md"), "w").write("\n".join(out) + "\n")


def kernels(tag, reps):
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    out = ["# %s — per-kernel metrics from `ncu --set full --clock-control none --import-source on`" % tag, ""]
    per_kernel = collections.OrderedDict()
    fresh = {}
    for rep in reps:
        # a .ncu-rep is exported here; a .csv is the same export done on the GPU box (a full capture can exceed what gpurun brings back)
        raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            name = short(r[hdr.index("Kernel Name")])
            d = {}
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    try:
                        v = float(r[i].replace(",", ""))
                    except ValueError:
                        continue
                    d[m] = v * UNIT.get(units[i], 1.0) if "bytes" in m else v
                    if m == "gpu__time_duration.sum":
                        d[m] = v * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(units[i], 1.0)
            d["grid"] = r[hdr.index("Grid Size")] if "Grid Size" in hdr else ""
            d["block"] = r[hdr.index("Block Size")] if "Block Size" in hdr else ""
            per_kernel.setdefault(name, []).append(d)
    for name, ds in per_kernel.items():
        out += ["## `%s` (%d captured launches, source: %s)" % (name, len(ds), ", ".join(os.path.basename(r) for r in reps)), "",
                "| launch | grid x block | duration us | dram read MB | dram write MB | dram % of peak | SM % | warps active % | regs | tensor pipe active % |", "|---|---|---:|---:|---:|---:|---:|---:|---:|---:|"]
        for i, d in enumerate(ds):
            out.append("| %d | %s x %s | %.1f | %.2f | %.2f | %.1f | %.1f | %.1f | %d | %.1f |" % (
                i, d.get("grid", ""), d.get("block", ""), d.get("gpu__time_duration.sum", 0), d.get("dram__bytes_read.sum", 0) / 1e6, d.get("dram__bytes_write.sum", 0) / 1e6,
                d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0), d.get("sm__throughput.avg.pct_of_peak_sustained_elapsed", 0),
                d.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0), int(d.get("launch__registers_per_thread", 0)),
                d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0)))
        out.append("")
        # traffic per launch: median over the largest-grid launches of this kernel
        big = max(ds, key=lambda d: d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0))
        key = name.split("<")[0]
        if "conv_sw_kernel" in key:                     # one template, three layers: tell them apart by the rows-per-item parameter
            key = {"441": "qnet_conv1", "100": "qnet_conv2", "81": "qnet_conv3"}.get(name.split("ConvGeom<")[1].split(",")[0].strip(), key)
        key = key.replace("qnet::", "")
        fresh.setdefault(key, 0.0)                      # several template shapes share a key: keep the biggest launch (the bench launch)
        fresh[key] = max(fresh[key], big.get("dram__bytes_read.sum", 0) + big.get("dram__bytes_write.sum", 0))
    traffic.update(fresh)
    open(os.path.join(ROOT, "profiles", tag + "_kernels.md"), "w").write("\n".join(out) + "\n")
    json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    tag = sys.argv[1]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    launches(tag, sys.argv[2])
    kernels(tag, sys.argv[3:])
    print(open(os.path.join(ROOT, "profiles", tag + "_launches.md")).read())
    print(open(os.path.join(ROOT, "profiles", tag + "_kernels.md")).read())
