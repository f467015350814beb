#!/usr/bin/env python
"""Accuracy-vs-iteration curves from training logs (the reference's only evaluation harness,
example/ResNet18/draw_curve.py): parses ' * All Loss ... Prec@1 ...' lines from text logs or
'acc1_val' records from train.jsonl files.   python draw_curve.py aps.log no_aps.log"""
import json
import re
import sys


def parse(path):
    xs, ys, step = [], [], 0
    for line in open(path):
        if path.endswith(".jsonl"):
            r = json.loads(line)
            if "acc1_val" in r:
                xs.append(r["step"]); ys.append(r["acc1_val"])
            continue
        m = re.search(r"Iter: \[(\d+)/", line)
        if m:
            step = int(m.group(1))
        m = re.search(r"\* All Loss [\d.]+ Prec@1 ([\d.]+)", line)
        if m:
            xs.append(step); ys.append(float(m.group(1)))
    return xs, ys


if __name__ == "__main__":
    import matplotlib
    matplotlib.use("Agg")
    import matplotlib.pyplot as plt
    for p in sys.argv[1:] or ["aps.log", "no_aps.log"]:
        x, y = parse(p)
        plt.plot(x, y, label=p)
    plt.xlabel("iteration"); plt.ylabel("Prec@1"); plt.legend(); plt.savefig("curve.png", dpi=150)
    print("wrote curve.png")
