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
    out = {"kernel": kernel, "report": a.rep, "samples": a.samples, "warp_instructions": warp_total,
           "thread_instructions_per_sample_full_warps": per(warp_total), "thread_instructions_per_sample_active_threads": thread_total / a.samples,
           "per_class": {k: per(v) for k, v in per_cls.most_common()}, "per_opcode": {k: per(v) for k, v in per_op.most_common()},
           "mufu_per_sample": per(sum(v for k, v in per_op.items() if k.startswith("MUFU"))),
           "note": "warp instructions x 32 / samples: an issue slot is spent per warp instruction whatever its active mask"}
    lines = [f"# Executed instruction mix — {kernel}", f"# {a.rep}: {warp_total:,} warp instructions for {a.samples:,.0f} pixel.light samples",
             "", f"issue slots (warp instr x 32) per sample: **{out['thread_instructions_per_sample_full_warps']:.1f}**; MUFU per sample: **{out['mufu_per_sample']:.2f}**", "",
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
