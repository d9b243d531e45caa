"""Turn an `ncu --csv` metrics log of tools/quick_perf.py into profiles/kernel_counts.json: warp-level instruction,
FP64-instruction and flop counts and DRAM bytes per track-step of the forward kernel (with and without the smoother
tape) and of the backward kernel.  bench.py reads that file for its roofline block (no constants in bench.py).

usage: python tools/ncu_counts.py <counts.csv> <tracks> <steps> [--full-cov counts_full.csv] [--note text] > profiles/kernel_counts.json
The capture: ncu --clock-control none --csv --metrics <METRICS below> -k regex:'ukf_forward|urtss_backward' python tools/quick_perf.py ..."""
import csv, io, json, sys

METRICS = ("smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,"
           "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,"
           "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,"
           "smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread")


def parse(path):
    """-> list of launches, each {kernel, metric: value} (ncu --csv long format, one row per metric)."""
    text = open(path).read()
    start = text.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    by_id = {}
    for r in rows:
        d = by_id.setdefault(r["ID"], {"kernel": r["Kernel Name"]})
        v, unit = r["Metric Value"].replace(",", ""), r.get("Metric Unit", "")
        try:
            v = float(v)
        except ValueError:
            continue
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0,
                 "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}.get(unit, 1.0)
        d[r["Metric Name"]] = v * scale
    return [by_id[k] for k in sorted(by_id, key=int)]


def per_step(launch, track_steps):
    g = launch.get
    warp = lambda name: (g(name) or 0.0) / 32.0 / (track_steps / 32.0)   # thread-level count -> per thread-step (= per track-step)
    dfma, dmul, dadd = (warp(f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum") for op in ("dfma", "dmul", "dadd"))
    out = {"inst": (g("smsp__inst_executed.sum") or 0.0) / (track_steps / 32.0),
           "fp64_instr": (g("sm__inst_executed_pipe_fp64.sum") or 0.0) / (track_steps / 32.0),
           "dfma": dfma, "dmul": dmul, "dadd": dadd, "flops": 2.0 * dfma + dmul + dadd,
           "dram_bytes": ((g("dram__bytes_read.sum") or 0.0) + (g("dram__bytes_write.sum") or 0.0)) / track_steps,
           "dram_read": (g("dram__bytes_read.sum") or 0.0) / track_steps, "dram_write": (g("dram__bytes_write.sum") or 0.0) / track_steps,
           "duration_ms_under_ncu": 1e3 * (g("gpu__time_duration.sum") or 0.0),
           "fp64_pipe_active_pct": g("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
           "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
           "registers": g("launch__registers_per_thread")}
    return out


def classify(launches, track_steps):
    """quick_perf launches: forward with tape (1 + reps), backward (1 + reps), forward without tape (1 + reps); the
    tape launches write 152 B per track-step more than the tape-less ones."""
    fwd = [l for l in launches if "ukf_forward" in l["kernel"]]
    bwd = [l for l in launches if "urtss_backward" in l["kernel"]]
    half = len(fwd) // 2
    return {"forward": per_step(fwd[half - 1], track_steps), "forward_no_tape": per_step(fwd[-1], track_steps),
            "backward": per_step(bwd[-1], track_steps)}


if __name__ == "__main__":
    path, tracks, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    out = classify(parse(path), tracks * steps)
    if "--full-cov" in sys.argv:
        full = classify(parse(sys.argv[sys.argv.index("--full-cov") + 1]), tracks * steps)
        for k in out:
            out[k]["dram_bytes_full_cov"] = full[k]["dram_bytes"]
    if "--rf-model" in sys.argv:   # cycles per warp-step from tools/fp64_operand_model.py: forward=..,forward_no_tape=..,backward=..
        for item in sys.argv[sys.argv.index("--rf-model") + 1].split(","):
            k, v = item.split("=")
            out[k]["fp64_rf_cycles"] = float(v)
    note = sys.argv[sys.argv.index("--note") + 1] if "--note" in sys.argv else ""
    out["source"] = (f"ncu --metrics ... --clock-control none over tools/quick_perf.py --tracks {tracks} --steps {steps} --packed "
                     f"(tools/ncu_counts.py); per track-step = per launch / ({tracks} x {steps}). {note}").strip()
    json.dump(out, sys.stdout, indent=1)
    print()
