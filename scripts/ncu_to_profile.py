#!/usr/bin/env python
"""Turn an `ncu --set full` report (gpurun_out/*.ncu-rep) into a small text summary under profiles/.

    python scripts/ncu_to_profile.py gpurun_out/prof_x.ncu-rep profiles/r01_x.md "title" [bench-log]

Reads the report with `ncu -i ... --page raw --csv` and `--page source --csv --print-source sass`
(works on the builder box: no GPU needed).  The bench log (the plain, un-profiled run of the same
command) is quoted so the CUDA-event time stands next to the ncu (cold-cache, serialised) time.
"""
import collections
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__cycles_elapsed.max",
]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True, check=True).stdout


def raw_metrics(rep):
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, r):
            d[h] = (v, u)
        out.append(d)
    return out


def sass(rep):
    txt = ncu(rep, "--page", "source", "--csv", "--print-source", "sass")
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, data = None, []
    for r in rows:
        if r and r[0] == "Kernel Name":
            if data:
                break
            continue
        if r and r[0] == "Address":
            hdr = r
        elif hdr and len(r) == len(hdr):
            data.append(r)
    return hdr, data


def main():
    rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
    log = sys.argv[4] if len(sys.argv) > 4 else None
    lines = [f"# {title}", "", f"source report: `{rep}` (scratch, not committed); capture: "
             "`ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 1`", ""]
    if log:
        try:
            last = [l for l in open(log) if l.startswith("{")][-1]
            j = json.loads(last)
            lines += ["## un-profiled run of the same command (CUDA events)", "",
                      f"- value: {j['value']:.4g} {j['unit']}, ms_per_step {j['ms_per_step']:.5f}",
                      f"- workload: {j['config']['workload']}",
                      f"- clocks: {j.get('clocks')}", ""]
        except Exception as e:  # noqa
            lines += [f"(bench log {log} unreadable: {e})", ""]
    for i, m in enumerate(raw_metrics(rep)):
        lines += [f"## launch {i}: {m.get('Kernel Name', ('?', ''))[0][:100]}", ""]
        for k in KEYS:
            if k in m:
                lines.append(f"- {k}: {m[k][0]} {m[k][1]}")
        lines.append("")
    hdr, data = sass(rep)
    if hdr:
        ia, isrc, ismp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
        tot = sum(int(r[ia]) for r in data) or 1
        ops = collections.Counter()
        for r in data:
            t = r[isrc].split()
            op = t[1] if t[0].startswith("@") else t[0]
            ops[op.split(".")[0]] += int(r[ia])
        lines += ["## SASS opcode mix (warp instructions executed)", "",
                  ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in ops.most_common(20)), ""]
        stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        agg = {hdr[i]: sum(int(r[i] or 0) for r in data) for i in stall_cols}
        ts = sum(int(r[ismp]) for r in data) or 1
        lines += ["## warp-stall samples (all)", "",
                  ", ".join(f"{k} {100 * v / ts:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]), "",
                  "## 15 most-sampled SASS instructions", ""]
        idx = sorted(range(len(data)), key=lambda i: -int(data[i][ismp]))[:15]
        for i in sorted(idx):
            r = data[i]
            st = sorted(((hdr[j], int(r[j] or 0)) for j in stall_cols), key=lambda kv: -kv[1])[:2]
            lines.append(f"- [{i}] samples {r[ismp]}, executed {r[ia]}: `{r[isrc].strip()[:60]}` {st}")
        lines.append("")
    open(out, "w").write("\n".join(lines))
    print("wrote", out)


if __name__ == "__main__":
    main()
