"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of the step)."""
import collections
import csv
import re
import sys


def summarize(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    hdr = rows[hi]
    kn, mv = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.defaultdict(list)
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        name = re.sub(r'\(.*', '', r[kn]).split('<')[0].replace('mli::', '').replace('void ', '')
        name = name.replace('unnamed>::', '')
        agg[name].append(float(r[mv].replace(',', '')))
    tot = sum(sum(v) for v in agg.values())
    out = ["| kernel | launches | avg us | share |", "|---|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| {k} | {len(v)} | {sum(v) / len(v) / 1e3:.2f} | {100 * sum(v) / tot:.1f}% |")
    out.append(f"\nTotal {tot / 1e3:.0f} us over {sum(len(v) for v in agg.values())} launches.")
    return "\n".join(out)


if __name__ == "__main__":
    print(summarize(sys.argv[1]))
