"""summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: the kernels of the last train step (delimited by the optimizer kernel)"""
import csv
import re
import sys


def main(path, marker="k_optim_multi"):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = [(int(r["ID"]), r["Kernel Name"], float(r["Metric Value"])) for r in csv.DictReader(lines)]
    idx = [i for i, (_, b, _) in enumerate(rows) if marker in b]
    if len(idx) < 2:
        print("no full step found (%d launches)" % len(rows)); return
    seg = rows[idx[-2] + 1: idx[-1] + 1]
    tot = sum(c for _, _, c in seg)
    for _, b, c in seg:
        print("%-64s %8.2f us %5.1f%%" % (re.sub(r"\(.*", "", b)[:64], c / 1e3, 100 * c / tot))
    print("step total %.1f us in %d launches (cold-cache, serialised: compare shares)" % (tot / 1e3, len(seg)))


if __name__ == "__main__":
    main(*sys.argv[1:])
