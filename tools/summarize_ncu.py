#!/usr/bin/env python3
"""Turn ncu output brought back in gpurun_out/ into the small text summaries committed under profiles/.

  summarize_ncu.py launches <launches.csv> <out.md>     per-kernel launch count / total time / share of the step
  summarize_ncu.py report   <file.ncu-rep> <out.md>     key metrics of one `ncu --set full` capture

Only reads files; the ncu CLI in this image does the .ncu-rep decoding (`ncu -i ... --page raw --csv`).
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEY_METRICS = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio",
    "local_load_requests", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("tvm::", "")


def launches(src, dst):
    rows = []
    with open(src, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r["Metric Name"] == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            if r["Metric Unit"] == "us":
                v *= 1e3
            elif r["Metric Unit"] == "ms":
                v *= 1e6
            rows.append((short(r["Kernel Name"]), v))
    agg = OrderedDict()
    for k, v in rows:
        c, t = agg.get(k, (0, 0.0))
        agg[k] = (c + 1, t + v)
    total = sum(t for _, t in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary\n\nsource: `{src}` ({len(rows)} launches captured, {total/1e6:.1f} ms summed kernel time; "
                "times are ncu's serialised cold-cache per-launch durations — use the SHARE, not the absolute)\n\n")
        f.write("| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {c} | {t/1e6:.2f} | {t/c/1e3:.1f} | {100*t/total:.1f}% |\n")


def report(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    lines = [l for l in out.splitlines(True) if l.startswith('"')]
    rd = list(csv.reader(io.StringIO("".join(lines))))
    header, units = rd[0], rd[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary\n\nsource: `{src}` (kept in gpurun_out/, not tracked)\n\n")
        for row in rd[2:]:
            d = dict(zip(header, row))
            u = dict(zip(header, units))
            f.write(f"## `{short(d.get('Kernel Name', '?'))}`  grid {d.get('Grid Size')} block {d.get('Block Size')}\n\n")
            f.write("| metric | value | unit |\n|---|---:|---|\n")
            for m in KEY_METRICS:
                if m in d and d[m] != "":
                    f.write(f"| {m} | {d[m]} | {u.get(m, '')} |\n")
            f.write("\n")


def table(src, dst):
    """one row per kernel name of an all-kernels `--set full` capture: time, achieved DRAM GB/s, issue utilisation"""
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    lines = [l for l in out.splitlines(True) if l.startswith('"')]
    rd = list(csv.reader(io.StringIO("".join(lines))))
    header, units = rd[0], rd[1]
    u = dict(zip(header, units))

    def num(d, k, default=0.0):
        try:
            return float(d.get(k, "").replace(",", ""))
        except ValueError:
            return default

    def to_bytes(v, unit):
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)

    def to_ns(v, unit):
        return v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(unit, 1)

    agg = OrderedDict()
    for row in rd[2:]:
        d = dict(zip(header, row))
        name = short(d.get("Kernel Name", "?"))
        if name.startswith("air_chunk_"):
            name = "air_chunk_* (generated AIR kernels)"
        a = agg.setdefault(name, dict(n=0, ns=0.0, rd=0.0, wr=0.0, issue=0.0, alu=0.0, fma=0.0, occ=0.0, regs=0))
        t = to_ns(num(d, "gpu__time_duration.sum"), u.get("gpu__time_duration.sum", "ns"))
        a["n"] += 1; a["ns"] += t
        a["rd"] += to_bytes(num(d, "dram__bytes_read.sum"), u.get("dram__bytes_read.sum", "byte"))
        a["wr"] += to_bytes(num(d, "dram__bytes_write.sum"), u.get("dram__bytes_write.sum", "byte"))
        a["issue"] += t * num(d, "smsp__issue_active.avg.pct_of_peak_sustained_active")
        a["alu"] += t * num(d, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active")
        a["fma"] += t * num(d, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active")
        a["occ"] += t * num(d, "sm__warps_active.avg.pct_of_peak_sustained_active")
        a["regs"] = max(a["regs"], int(num(d, "launch__registers_per_thread")))
    total = sum(a["ns"] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# every kernel of one prove(), `ncu --set full`\n\nsource: `{src}` (kept in gpurun_out/, not tracked); "
                "durations are ncu's serialised cold-cache per-launch times; pipe / issue figures are time-weighted means\n\n")
        f.write("| kernel | launches | ms | share | DRAM GB/s (rd+wr) | issue % | ALU pipe % | FMA pipe % | warps active % | regs |\n"
                "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
            t = a["ns"] or 1.0
            f.write(f"| `{k}` | {a['n']} | {a['ns']/1e6:.3f} | {100*a['ns']/total:.1f}% | {(a['rd']+a['wr'])/t:.0f} | "
                    f"{a['issue']/t:.0f} | {a['alu']/t:.0f} | {a['fma']/t:.0f} | {a['occ']/t:.0f} | {a['regs']} |\n")


def table_long(src, dst):
    """like `table`, from the long-format CSV that `ncu --metrics ... --csv --log-file` writes (one row per launch and metric)"""
    with open(src, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    per = OrderedDict()
    for r in csv.DictReader(io.StringIO("".join(lines))):
        d = per.setdefault(r["ID"], {"name": short(r["Kernel Name"])})
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = r["Metric Unit"]
        v *= {"us": 1e3, "ms": 1e6, "s": 1e9, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        d[r["Metric Name"]] = v
    agg = OrderedDict()
    for d in per.values():
        name = "air_chunk_* (generated AIR kernels)" if d["name"].startswith("air_chunk_") else d["name"]
        a = agg.setdefault(name, dict(n=0, ns=0.0, by=0.0, issue=0.0, alu=0.0, fma=0.0, occ=0.0, regs=0))
        t = d.get("gpu__time_duration.sum", 0.0)
        a["n"] += 1; a["ns"] += t
        a["by"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        a["issue"] += t * d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0)
        a["alu"] += t * d.get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 0.0)
        a["fma"] += t * d.get("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 0.0)
        a["occ"] += t * d.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0.0)
        a["regs"] = max(a["regs"], int(d.get("launch__registers_per_thread", 0)))
    total = sum(a["ns"] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# every kernel of one prove()\n\nsource: `{src}` ({len(per)} launches; ncu's serialised cold-cache per-launch durations; "
                "pipe / issue figures are time-weighted means over the launches of a kernel)\n\n")
        f.write("| kernel | launches | ms | share | DRAM GB/s (rd+wr) | issue % | ALU pipe % | FMA pipe % | warps active % | regs |\n"
                "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
            t = a["ns"] or 1.0
            f.write(f"| `{k}` | {a['n']} | {a['ns']/1e6:.3f} | {100*a['ns']/total:.1f}% | {a['by']/t:.0f} | "
                    f"{a['issue']/t:.0f} | {a['alu']/t:.0f} | {a['fma']/t:.0f} | {a['occ']/t:.0f} | {a['regs']} |\n")


if __name__ == "__main__":
    {"launches": launches, "report": report, "table": table, "table_long": table_long}[sys.argv[1]](sys.argv[2], sys.argv[3])
