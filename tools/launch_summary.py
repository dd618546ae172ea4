"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time and
share of one denoise step (delimited by the time-embedding cos kernel)."""
import collections
import csv
import re
import sys


def main(path, which=-1):
    rows = list(csv.reader(open(path, errors="ignore")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    idx = {h: i for i, h in enumerate(hdr)}
    L = []
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        try:
            v = float(r[idx["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        unit = r[idx["Metric Unit"]]
        v = v / 1e3 if unit.startswith("n") else (v * 1e3 if unit.startswith("m") else v)
        L.append((r[idx["Kernel Name"]], v, r[idx["Grid Size"]], r[idx["Block Size"]]))
    cos = [i for i, (n, _, _, _) in enumerate(L) if "cos_kernel" in n]
    bounds = cos + [len(L)]
    which = which if which >= 0 else len(cos) + which
    step = L[bounds[which]:bounds[which + 1]]
    tot = sum(v for _, v, _, _ in step)
    print("total launches in file: %d; UNet forwards found: %d; summarising forward #%d: %d launches, %.1f us"
          % (len(L), len(cos), which, len(step), tot))
    OURS = ("fresco", "kv_compact", "temporal_attn", "warp_chain", "warp_blend", "flow_warp", "adain", "adam_kernel",
            "gram_", "mapping_single", "warp_loss", "gmflow_corr")

    def short(n):
        n = re.sub(r"\(.*", "", n).replace("void ", "")
        return n[:72]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, v, g, b in step:
        k = short(n)
        if any(s in n for s in OURS):
            k = "[ours] " + k
        agg[k][0] += 1
        agg[k][1] += v
    print("%-82s %5s %12s %7s" % ("kernel", "n", "time_us", "share"))
    for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1])[:30]:
        print("%-82s %5d %12.1f %6.1f%%" % (k, c, v, 100 * v / tot))
    ours = sum(v for k, (c, v) in agg.items() if k.startswith("[ours]"))
    n_ours = sum(c for k, (c, v) in agg.items() if k.startswith("[ours]"))
    print("ours: %d launches, %.1f us, %.1f%% of the step" % (n_ours, ours, 100 * ours / tot))
    print("\nper-launch list of our kernels:")
    for n, v, g, b in step:
        if any(s in n for s in OURS):
            print("   %-64s grid=%-16s block=%-12s %9.1f us" % (short(n), g, b, v))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else -1)
