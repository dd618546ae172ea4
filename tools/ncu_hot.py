"""Hot spots of one kernel from an `ncu --import-source on` report:   python tools/ncu_hot.py report.ncu-rep [top]
Per SASS instruction: executed warp-instructions (share of the kernel's) and stall samples by reason; then totals by
opcode, so polling loops and the real arithmetic can be told apart."""
import collections
import csv
import io
import re
import subprocess
import sys


def main(path, top=40):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    recs = []
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or r[0] == "Address":
            if r and r[0] == "Kernel Name":
                break
            continue
        ex = int(r[col["Instructions Executed"]] or 0)
        smp = int(r[col["# Samples"]] or 0)
        st = {h[6:]: int(r[col[h]] or 0) for h in stall_cols}
        recs.append((r[col["Address"]], r[col["Source"]].strip(), ex, smp, st))
    tot_ex = sum(x[2] for x in recs) or 1
    tot_s = sum(x[3] for x in recs) or 1
    print("# %s : %d SASS instructions, %d warp-instructions executed, %d stall samples" % (path, len(recs), tot_ex, tot_s))
    print("## by executed count")
    for i, (a, src, ex, smp, st) in sorted(enumerate(recs), key=lambda t: -t[1][2])[:top]:
        top_st = ",".join("%s=%d" % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:3] if kv[1])
        print("%5d %6.2f%% ex %6.2f%% smp  %-70s %s" % (i, 100.0 * ex / tot_ex, 100.0 * smp / tot_s, src[:70], top_st))
    print("## by stall samples")
    for i, (a, src, ex, smp, st) in sorted(enumerate(recs), key=lambda t: -t[1][3])[:top]:
        top_st = ",".join("%s=%d" % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:3] if kv[1])
        print("%5d %6.2f%% ex %6.2f%% smp  %-70s %s" % (i, 100.0 * ex / tot_ex, 100.0 * smp / tot_s, src[:70], top_st))
    by_op = collections.Counter()
    for a, src, ex, smp, st in recs:
        op = re.sub(r"^@!?U?P\w+\s+", "", src).split(" ")[0].split(".")[0]
        by_op[op] += ex
    print("## executed by opcode")
    print("  ".join("%s %.1f%%" % (op, 100.0 * n / tot_ex) for op, n in by_op.most_common(24)))
    agg = collections.Counter()
    for a, src, ex, smp, st in recs:
        for k, v in st.items():
            agg[k] += v
    print("## stall samples by reason")
    print("  ".join("%s %.1f%%" % (k, 100.0 * v / tot_s) for k, v in agg.most_common(12)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
