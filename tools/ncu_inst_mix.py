#!/usr/bin/env python
"""Executed-instruction mix of one captured kernel from an .ncu-rep (ncu --set full --import-source on): per SASS opcode
and per pipe class, as thread instructions per pixel.light sample.  This is the table the issue-slot roof of bench.py's
`config3` key is computed from (profiles/r02_inst_mix_*.json).

    python tools/ncu_inst_mix.py REP --samples 1073741824 [--json OUT.json] [--md OUT.md]
"""
import argparse
import collections
import csv
import io
import json
import subprocess

from sass_stats import classify


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--samples", type=float, required=True, help="pixel.light samples the captured launch processed")
    ap.add_argument("--json")
    ap.add_argument("--md")
    ap.add_argument("--top", type=int, default=28)
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    kernel = rows[0][1]
    hdr = rows[1]
    i_src, i_ex, i_th, i_samp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    per_op, per_cls, stall = collections.Counter(), collections.Counter(), collections.Counter()
    warp_total = thread_total = samples_total = 0
    for r in rows[2:]:
        if len(r) <= i_th:
            continue
        toks = r[i_src].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        op = op.rstrip(";")
        w, t, s = int(r[i_ex] or 0), int(r[i_th] or 0), int(r[i_samp] or 0)
        base = op.split(".")[0]
        key = op if base == "MUFU" else base
        per_op[key] += w
        per_cls[classify(op)] += w
        stall[key] += s
        warp_total += w
        thread_total += t
        samples_total += s
    per = lambda w: w * 32.0 / a.samples
    # the un-instrumented timing pass of the same capture (raw page): what the issue-slot roof is computed from — the
    # source-page counters come from an instrumented replay whose barrier-polling loops spin longer
    raw_rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True,
                                                         check=True).stdout)))
    rh, rd = raw_rows[0], raw_rows[2]
    rawv = lambda k: float(rd[rh.index(k)]) if k in rh and rd[rh.index(k)] not in ("", "no data") else None
    timed = {k: rawv(k) for k in ("smsp__inst_executed.sum", "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
                                  "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.max.pct_of_peak_sustained_active",
                                  "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
                                  "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum")}
    out = {"kernel": kernel, "report": a.rep, "samples": a.samples, "warp_instructions": warp_total,
           "thread_instructions_per_sample_full_warps": per(warp_total), "thread_instructions_per_sample_active_threads": thread_total / a.samples,
           "per_class": {k: per(v) for k, v in per_cls.most_common()}, "per_opcode": {k: per(v) for k, v in per_op.most_common()},
           "mufu_per_sample": per(sum(v for k, v in per_op.items() if k.startswith("MUFU"))),
           "timed_pass": timed,
           "issue_slots_per_sample": (timed["smsp__inst_executed.sum"] * 32.0 / a.samples) if timed.get("smsp__inst_executed.sum") else per(warp_total),
           "note": "warp instructions x 32 / samples: an issue slot is spent per warp instruction whatever its active mask"}
    lines = [f"# Executed instruction mix — {kernel}", f"# {a.rep}: {warp_total:,} warp instructions for {a.samples:,.0f} pixel.light samples",
             "", f"issue slots (warp instr x 32) per sample: **{out['issue_slots_per_sample']:.1f}** in the timing pass (smsp__inst_executed.sum), "
                 f"{out['thread_instructions_per_sample_full_warps']:.1f} in the instrumented source-counter pass; MUFU per sample: **{out['mufu_per_sample']:.2f}**", "",
             "| pipe class | thread instr / sample | share |", "|---|---|---|"]
    for k, v in per_cls.most_common():
        lines.append(f"| {k} | {per(v):.2f} | {100.0 * v / warp_total:.1f} % |")
    lines += ["", "| opcode | thread instr / sample | share of issue slots | share of warp-stall samples |", "|---|---|---|---|"]
    for k, v in per_op.most_common(a.top):
        lines.append(f"| {k} | {per(v):.2f} | {100.0 * v / warp_total:.1f} % | {100.0 * stall[k] / max(samples_total, 1):.1f} % |")
    text = "\n".join(lines) + "\n"
    print(text)
    if a.json:
        json.dump(out, open(a.json, "w"), indent=1)
    if a.md:
        open(a.md, "w").write(text)


if __name__ == "__main__":
    main()
