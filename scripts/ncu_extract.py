"""Reads gpurun_out/*.ncu-rep captures here (no GPU needed) and writes the judged summaries under profiles/:
one <name>.summary.txt per capture (the metrics B200_PROFILING.md names) and profiles/ncu_gemm_traffic.json
(dram bytes per launch of the dominant GEMM, read by bench.py's roofline.traffic)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg", "lts__throughput.avg",
        "lts__t_bytes.sum", "launch__registers_per_thread", "launch__grid_size", "launch__cluster",
        "sm__cycles_elapsed.max", "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime", "sm__mem_tensor_cycles_active.avg",
        "l1tex__data_pipe_tc", "sm__inst_executed.avg.per_cycle_elapsed")
SKIP = ("sm__ops_path", ".min", ".max.", ".sum.pct", "per_second")


def raw(path):
    r = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    if len(rows) < 3:
        raise SystemExit(f"{path}: no data ({r.stderr[:200]})")
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, row)) for row in rows[2:]], dict(zip(hdr, units))


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def main():
    out_traffic = {}
    for path in sys.argv[1:]:
        name = os.path.basename(path).replace(".ncu-rep", "")
        recs, units = raw(path)
        lines = []
        for rec in recs:
            lines.append(f"kernel: {rec.get('Kernel Name', '')[:120]}  grid {rec.get('Grid Size')} block {rec.get('Block Size')}")
            for k, v in rec.items():
                if any(key in k for key in KEYS) and not any(x in k for x in SKIP):
                    lines.append(f"  {k} [{units.get(k, '')}] = {v}")
            rd, wr = num(rec.get("dram__bytes_read.sum", "")), num(rec.get("dram__bytes_write.sum", ""))
            if rd is not None and wr is not None:
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                rd *= scale.get(units.get("dram__bytes_read.sum", "byte"), 1.0)
                wr *= scale.get(units.get("dram__bytes_write.sum", "byte"), 1.0)
                lines.append(f"  => DRAM traffic of this launch: {(rd + wr) / 1e6:.1f} MB (read {rd / 1e6:.1f}, write {wr / 1e6:.1f})")
                act = num(rec.get("TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", ""))
                el = num(rec.get("sm__cycles_elapsed.max", ""))
                if act and el:
                    lines.append(f"  => tensor pipe active {act / 4 / el * 100:.1f} % of the elapsed SM cycles (the realtime counter ticks once per "
                                 f"sub-partition: hmma_cycles_active_realtime / 4 / sm__cycles_elapsed)")
                out_traffic[name] = rd + wr
        with open(os.path.join(ROOT, "profiles", name + ".summary.txt"), "w") as f:
            f.write("\n".join(lines) + "\n")
        print("\n".join(lines[:60]))
    if out_traffic:
        path = os.path.join(ROOT, "profiles", "ncu_gemm_traffic.json")
        prev = {}
        if os.path.isfile(path):
            prev = json.load(open(path))
        prev.setdefault("captures", {}).update(out_traffic)
        # bench.py's roofline.traffic: the mean over the captured FFN GEMMs (same unit as achieved: per launch)
        vals = [v for k, v in prev["captures"].items() if "r2_ncu_fc" in k]
        if vals:
            prev["bf16x3"] = sum(vals) / len(vals)
        json.dump(prev, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
