#!/usr/bin/env python
"""Pretty-print selected keys of a bench.py JSON line: python tools/show_bench.py FILE [key ...]"""
import json
import sys


def show(k, v, ind=0):
    if isinstance(v, dict):
        print(" " * ind + k + ":")
        for kk, vv in v.items():
            show(kk, vv, ind + 2)
    else:
        print(" " * ind + f"{k}: {str(v)[:120]}")


d = json.loads(open(sys.argv[1]).readline())
keys = sys.argv[2:] or [k for k in d if k != "config"]
for k in keys:
    show(k, d.get(k))
