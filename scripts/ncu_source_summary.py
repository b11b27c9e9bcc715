"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source sass` output: per kernel, the share of warp-stall samples and of
executed warp instructions in windows of consecutive SASS instructions, with the dominant opcodes and the lane utilisation.
    python scripts/ncu_source_summary.py source.csv [kernel-substring] [window]"""
import csv
import sys


def main():
    path = sys.argv[1]
    pick = sys.argv[2] if len(sys.argv) > 2 else ""
    win = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    rows = list(csv.reader(open(path)))
    k = 0
    while k < len(rows):
        if rows[k] and rows[k][0] == "Kernel Name":
            name = rows[k][1]
            hdr = rows[k + 1]
            j = k + 2
            while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
                j += 1
            data = [r for r in rows[k + 2:j] if len(r) == len(hdr)]
            k = j
            if pick not in name:
                continue
            iS, iI, iT, iSrc = (hdr.index(c) for c in ("# Samples", "Instructions Executed", "Thread Instructions Executed", "Source"))
            tot_s = sum(int(r[iS]) for r in data) or 1
            tot_i = sum(int(r[iI]) for r in data) or 1
            tot_t = sum(int(r[iT]) for r in data)
            print(f"== {name[:100]}\n   {len(data)} SASS instructions, {tot_s} samples, {tot_i} warp instructions, lanes/instr {tot_t / tot_i:.1f}")
            for a in range(0, len(data), win):
                w = data[a:a + win]
                s = sum(int(r[iS]) for r in w)
                i = sum(int(r[iI]) for r in w)
                t = sum(int(r[iT]) for r in w)
                ops = {}
                for r in w:
                    tk = r[iSrc].split()
                    op = (tk[1] if tk[0].startswith("@") else tk[0]).split(".")[0]
                    ops[op] = ops.get(op, 0) + int(r[iI])
                top = ", ".join(f"{o} {100 * c / max(i, 1):.0f}%" for o, c in sorted(ops.items(), key=lambda x: -x[1])[:5])
                print(f"   [{a:5d}] samples {100 * s / tot_s:5.1f}%  instr {100 * i / tot_i:5.1f}%  lanes {t / max(i, 1):4.1f}  {top}")
        else:
            k += 1


if __name__ == "__main__":
    main()
