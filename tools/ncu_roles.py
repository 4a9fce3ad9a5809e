"""Where a warp-specialised kernel's warps spend their time, from an ncu report taken with
--set full --import-source on:  python tools/ncu_roles.py REPORT.ncu-rep KERNEL_REGEX [launch_skip] [top]

ncu's source page gives stall samples per SASS instruction.  Warp roles (TMA producer, MMA issuer, epilogue,
converter ...) execute disjoint code, and within a role every instruction of the per-tile loop is executed
the same number of times, so grouping consecutive instructions by their "Instructions Executed" count
separates the roles; the samples on the mbarrier try-wait loops against the role's total say how long the
role waits for its neighbours.  Prints (1) the regions with their sample share and dominant opcodes and
(2) the individual instructions with the most samples."""
import collections
import csv
import io
import subprocess
import sys


def source_rows(report, kernel, skip):
    cmd = ["ncu", "-i", report, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel,
           "--launch-skip", str(skip), "--launch-count", "1"]
    text = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    header = next(r for r in rows if "Source" in r and "# Samples" in r)
    idx = {n: i for i, n in enumerate(header)}
    seen, out = set(), []
    for r in rows:
        if len(r) == len(header) and r[0].startswith("0x") and r[0] not in seen:     # the page repeats per view
            seen.add(r[0])
            out.append({"addr": r[0], "sass": r[idx["Source"]].strip(), "samples": int(r[idx["# Samples"]] or 0),
                        "executed": int(r[idx["Instructions Executed"]] or 0)})
    return out


def opcode(sass):
    parts = sass.split()
    return parts[1] if parts and parts[0].startswith("@") and len(parts) > 1 else (parts[0] if parts else "")


def main():
    if len(sys.argv) < 3:
        raise SystemExit(__doc__)
    report, kernel = sys.argv[1], sys.argv[2]
    skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    rows = source_rows(report, kernel, skip)
    total = sum(r["samples"] for r in rows) or 1
    print("%d instructions, %d samples" % (len(rows), total))
    # regions: runs of instructions whose execution counts are within 2x of the run's first hot instruction
    regions, cur = [], None
    for i, r in enumerate(rows):
        ex = r["executed"]
        if cur and ex and cur["ref"] and 0.5 <= ex / cur["ref"] <= 2.0 or (cur and not ex):
            cur["rows"].append(r)
        else:
            cur = {"start": i, "ref": ex, "rows": [r]}
            regions.append(cur)
    print("\nregions (instruction range, executions per instruction, share of samples, dominant opcodes):")
    for g in regions:
        s = sum(r["samples"] for r in g["rows"])
        if s < total * 0.005:
            continue
        ops = collections.Counter(opcode(r["sass"]) for r in g["rows"]).most_common(4)
        waits = sum(r["samples"] for r in g["rows"] if "TRYWAIT" in r["sass"] or (opcode(r["sass"]) == "BRA" and r["samples"] > 0.3 * s))
        print("  [%5d, %5d)  x%-9d %5.1f%%  (%.0f%% of it on wait loops)  %s" %
              (g["start"], g["start"] + len(g["rows"]), g["ref"], 100.0 * s / total, 100.0 * waits / max(s, 1),
               ", ".join("%s x%d" % o for o in ops)))
    print("\ntop instructions:")
    for i, r in sorted(enumerate(rows), key=lambda t: -t[1]["samples"])[:top]:
        print("  %5d  %5.1f%%  x%-9d %s" % (i, 100.0 * r["samples"] / total, r["executed"], r["sass"][:100]))


if __name__ == "__main__":
    main()
