#!/usr/bin/env python3
"""Summarises ncu outputs brought back in gpurun_out/ into profiles/ (run here, no GPU needed).
  launches <launches.csv> <out.txt> <note>      per-kernel device-time shares
  full <raw.csv from `ncu -i X.ncu-rep --page raw --csv`> <out.json> <units_in_launch> <note>
"""
import collections, csv, io, json, sys


def launches(path, out, note):
    txt = open(path).read()
    txt = txt[txt.index('"ID"'):]
    rows = list(csv.DictReader(io.StringIO(txt)))
    agg = collections.OrderedDict()
    for r in rows:
        k = r["Kernel Name"].split("(")[0]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    lines = [f"{t / 1e6:10.3f} ms  {100 * t / tot:5.1f}%  x{n:4d}  {k}" for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])]
    open(out, "w").write(note + "\n(ncu per-launch times are cold-cache and serialised: shares only)\n" + "\n".join(lines) + f"\n{tot / 1e6:10.3f} ms  total\n")
    print("\n".join(lines))


KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
MULT = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1, "Tbyte": 1e12}


def full(path, out, units, note):
    rows = list(csv.reader(open(path)))
    res = []
    for vals in rows[2:]:
        d = dict(zip(rows[0], vals)); u = dict(zip(rows[0], rows[1]))
        tr = sum(float(d[k].replace(",", "")) * MULT[u[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        res.append({"kernel": d.get("Kernel Name"), "dram_bytes_per_launch": tr, "units_in_launch": units, "dram_bytes_per_unit": tr / units if units else None,
                    "metrics": {k: [d[k], u[k]] for k in KEYS if k in d}})
    json.dump({"note": note, "launches": res}, open(out, "w"), indent=1)
    for r in res:
        print(r["kernel"], r["dram_bytes_per_launch"], r["metrics"].get("sm__throughput.avg.pct_of_peak_sustained_elapsed"))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        full(sys.argv[2], sys.argv[3], float(sys.argv[4]), sys.argv[5])
