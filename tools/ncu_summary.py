"""One line per captured launch with the metrics the profiles/ summaries carry, from an `ncu --set full` report:
  python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rNN_ncu_x_summary.csv
(needs only ncu's own `--page raw --csv` export; runs in the CPU container)."""
import csv
import io
import subprocess
import sys

COLS = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "gpc__cycles_elapsed.avg.per_second", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__block_size",
        "launch__cluster_size"]
# units ncu may pick per report -> the unit the summaries use (ms, GHz, GB)
SCALE = {"gpu__time_duration.sum": {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3},
         "gpc__cycles_elapsed.avg.per_second": {"hz": 1e-9, "khz": 1e-6, "mhz": 1e-3, "ghz": 1.0, "cycle/nsecond": 1.0, "cycle/usecond": 1e-3,
                                                 "cycle/second": 1e-9},
         "dram__bytes_read.sum": {"byte": 1e-9, "kbyte": 1e-6, "mbyte": 1e-3, "gbyte": 1.0},
         "dram__bytes_write.sum": {"byte": 1e-9, "kbyte": 1e-6, "mbyte": 1e-3, "gbyte": 1.0}}


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, data = rows[0], rows[1], rows[2:]
    idx = {name: i for i, name in enumerate(head)}
    out = csv.writer(sys.stdout)
    out.writerow(COLS)
    for r in data:
        line = []
        for c in COLS:
            if c not in idx:
                line.append("")
                continue
            v = r[idx[c]]
            if c in SCALE:
                unit = units[idx[c]].strip().lower()
                if unit not in SCALE[c]:
                    raise SystemExit("unit %r of %s is not in the table" % (unit, c))
                v = "%.6f" % (float(v.replace(",", "")) * SCALE[c][unit])
            line.append(v)
        out.writerow(line)


if __name__ == "__main__":
    main()
