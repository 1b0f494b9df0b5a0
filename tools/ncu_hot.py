"""Hot spots of one kernel in an ncu report captured with --import-source on:
`python tools/ncu_hot.py report.ncu-rep KERNEL_REGEX [TOP]` prints the SASS instructions with the most warp-stall
samples (and their executed counts), plus totals per opcode."""
import csv
import subprocess
import sys
from collections import Counter


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-kernel-base", "function", "--kernel-name", f"regex:{rx}"], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for row in csv.reader(raw.splitlines()):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": [], "hdr": None}
            blocks.append(cur)
        elif cur is not None and row:
            if cur["hdr"] is None:
                cur["hdr"] = row
            else:
                cur["rows"].append(row)
    for b in blocks:
        h = b["hdr"]
        i_src, i_smp, i_exe = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
        rows = [(int(r[i_smp] or 0), int(r[i_exe] or 0), k, r[i_src].strip()) for k, r in enumerate(b["rows"])]
        tot_s, tot_e = sum(r[0] for r in rows), sum(r[1] for r in rows)
        print(f"=== {b['name']}: {len(rows)} SASS instructions, {tot_e} warp-instructions executed, {tot_s} samples")
        for s, e, k, src in sorted(rows, reverse=True)[:top]:
            print(f"  {100 * s / max(tot_s, 1):5.1f}% samples  {100 * e / max(tot_e, 1):5.1f}% exec  #{k:4d}  {src[:90]}")
        ops = Counter()
        for s, e, k, src in rows:
            op = src.split()[1] if src.startswith("@") else src.split()[0]
            ops[op.split(".")[0]] += e
        print("  executed by opcode:", ", ".join(f"{o} {100 * c / max(tot_e, 1):.1f}%" for o, c in ops.most_common(14)))


if __name__ == "__main__":
    main()
