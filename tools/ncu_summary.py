"""Turn ncu CSV exports into the tracked evidence under profiles/ (run here, no GPU needed).

  python tools/ncu_summary.py --raw gpurun_out/<x>_raw.csv --evals "<7 numbers>" --out profiles/<name>
      per-kernel summary (<name>_summary.txt), the selected counters as CSV (<name>_metrics.csv) and JSON
      (<name>.json: per constant FP64-pipe %, issue %, executed FP64 flops per integrand evaluation, stalls)
      from an `ncu -i <rep> --page raw --csv` export of the seven K2 launches + K1 at the headline size.
      --evals: integrand evaluations per constant of that run (tools/profile_case.py prints them), order
      NSp NSm NS_nfl VGG AGG VGG_nfl AGG_nfl.
  python tools/ncu_summary.py --calibration gpurun_out/<cal>.csv --npts 1048576 --out profiles/k2_flops_per_eval.json
      algorithmic flops per integrand evaluation per constant from the counter run of tools/calibrate_k2_flops.py
      (merged into the JSON next to the executed numbers of the full-size capture, if --executed-json is given).
"""
import argparse
import csv
import json
import re
import sys

KID = {7: ("KV1_NSp", 0), 8: ("KV1_NSm", 1), 9: ("KV1_NS_nfl", 2), 18: ("KV1_GG", 3), 19: ("KA1_GG", 4),
       20: ("KV1_GG_nfl", 5), 21: ("KA1_GG_nfl", 6)}
KEEP = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "sm__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__icc_request_hit_rate.pct",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
]
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "Tbyte": 1e12}


def fnum(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return None


def pair_of(name):
    """the two constants of an assemble_cst_pair_kernel<KA, KB, ...> launch (one launch computes both), else None"""
    m = re.search(r"assemble_cst_pair_kernel<\(?(?:int\))?(\d+), \(?(?:int\))?(\d+)", name)
    return (int(m.group(1)), int(m.group(2))) if m else None


def kid_of(name):
    m = re.search(r"assemble_cst_\w+_kernel<\(?(?:int\))?(\d+)", name)
    return int(m.group(1)) if m else None


def read_raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def thread_inst(row, idx, hdr, op):
    """thread-level dadd/dmul/dfma count: the .sum metric if present, else per_cycle_elapsed x elapsed cycles"""
    k = "smsp__sass_thread_inst_executed_op_%s_pred_on.sum" % op
    if k in idx and fnum(row[idx[k]]) is not None:
        return fnum(row[idx[k]])
    k2 = k + ".per_cycle_elapsed"
    cyc = fnum(row[idx["smsp__cycles_elapsed.avg"]]) if "smsp__cycles_elapsed.avg" in idx else fnum(row[idx["sm__cycles_elapsed.max"]])
    return fnum(row[idx[k2]]) * cyc


def summarize(args):
    hdr, units, rows = read_raw(args.raw)
    idx = {h: i for i, h in enumerate(hdr)}
    evals = [float(v) for v in args.evals.split()] if args.evals else None
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    out_txt, out_json, out_rows = [], {"source": args.raw, "note": args.note or "", "kernels": []}, []
    tot_ms = 0.0
    wsum = 0.0
    for r in rows:
        name = r[idx["Kernel Name"]]
        kid = kid_of(name)
        pair = pair_of(name)
        label = KID[kid][0] if kid in KID else ("K1 assemble_fixed_kernel" if "assemble_fixed" in name else name[:60])
        if pair:
            label = "%s + %s (one launch)" % (KID[pair[0]][0], KID[pair[1]][0])
        ms = fnum(r[idx["gpu__time_duration.sum"]])
        if units[idx["gpu__time_duration.sum"]] in ("us", "usecond"):
            ms /= 1e3
        elif units[idx["gpu__time_duration.sum"]] in ("ns", "nsecond"):
            ms /= 1e6
        elif units[idx["gpu__time_duration.sum"]] in ("s", "second"):
            ms *= 1e3
        pipe = fnum(r[idx["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]])
        issue = fnum(r[idx["smsp__issue_active.avg.pct_of_peak_sustained_active"]])
        dadd, dmul, dfma = (thread_inst(r, idx, hdr, op) for op in ("dadd", "dmul", "dfma"))
        flops = dadd + dmul + 2.0 * dfma
        dram = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            dram += fnum(r[idx[k]]) * UNIT_SCALE.get(units[idx[k]], 1.0)
        st = sorted(((fnum(r[idx[h]]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                     for h in stalls), reverse=True)[:8]
        k = {"kernel": label, "kid": kid, "ms": ms, "fp64_pipe_active_pct": pipe, "issue_active_pct": issue,
             "thread_inst": {"dadd": dadd, "dmul": dmul, "dfma": dfma}, "executed_fp64_flops": flops,
             "registers": fnum(r[idx["launch__registers_per_thread"]]), "warps_active": fnum(r[idx["sm__warps_active.avg.per_cycle_active"]]),
             "lanes_per_inst": fnum(r[idx["smsp__thread_inst_executed_per_inst_executed.ratio"]]),
             "icache_hit_pct": fnum(r[idx["sm__icc_request_hit_rate.pct"]]) if "sm__icc_request_hit_rate.pct" in idx else None,
             "dram_bytes": dram, "stalls_warp_cycles_per_issue": {n: v for v, n in st}}
        if pair:
            k["pair"] = list(pair)
        if kid in KID and evals:
            # a paired launch serves the evaluations of both constants' QUADPACK paths (most of them shared)
            ne = sum(evals[KID[q][1]] for q in pair) if pair else evals[KID[kid][1]]
            k["integrand_evaluations"] = ne
            k["executed_flops_per_eval"] = flops / ne
        if kid in KID:
            tot_ms += ms
            wsum += ms * pipe
        out_json["kernels"].append(k)
        out_rows.append([label] + [r[idx[m]] if m in idx else "" for m in KEEP])
        out_txt.append("== %s  (%s)\n   %.3f ms, %d regs, %.1f warps/SM active, issue slots %.1f %%, FP64 pipe active %.1f %%, %.2f lanes/inst, i-cache hit %s %%\n"
                       "   thread instructions: dadd %.4g dmul %.4g dfma %.4g -> %.4g executed FP64 flop%s; DRAM %.3g MB\n   stalls (warp-cycles per issued instruction): %s\n"
                       % (label, name[:90], ms, k["registers"], k["warps_active"], issue, pipe, k["lanes_per_inst"],
                          ("%.1f" % k["icache_hit_pct"]) if k["icache_hit_pct"] is not None else "n/a", dadd, dmul, dfma, flops,
                          (" = %.0f per integrand evaluation" % k["executed_flops_per_eval"]) if "executed_flops_per_eval" in k else "",
                          dram / 1e6, ", ".join("%s %.2f" % (n, v) for v, n in st)))
    if tot_ms > 0:
        out_json["k2_sum_ms"] = tot_ms
        out_json["k2_fp64_pipe_active_pct_time_weighted"] = wsum / tot_ms
        out_txt.append("K2 (all constants): %.1f ms serialised under ncu, FP64 pipe active %.1f %% (time-weighted)\n" % (tot_ms, wsum / tot_ms))
    with open(args.out + "_summary.txt", "w") as f:
        f.write((args.note + "\n" if args.note else "") + "".join(out_txt))
    with open(args.out + "_metrics.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + KEEP)
        w.writerow([""] + [units[idx[m]] if m in idx else "" for m in KEEP])
        w.writerows(out_rows)
    with open(args.out + ".json", "w") as f:
        json.dump(out_json, f, indent=1)
    print("".join(out_txt))


def calibrate(args):
    rows = list(csv.reader(l for l in open(args.calibration) if not l.startswith("==")))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    per = {}
    for r in rows[1:]:
        if len(r) < len(hdr) or "eval_integrand_kernel" not in r[idx["Kernel Name"]]:
            continue
        m = re.search(r"eval_integrand_kernel<\(?(?:tkb::KernelId\))?(\d+)", r[idx["Kernel Name"]])
        kid = int(m.group(1))
        op = re.search(r"op_(dadd|dmul|dfma)_pred_on", r[idx["Metric Name"]]).group(1)
        per.setdefault(kid, {})[op] = fnum(r[idx["Metric Value"]])
    flops, counters = [0.0] * 7, {}
    for kid, (name, c) in KID.items():
        d = per[kid]
        flops[c] = (d["dadd"] + d["dmul"] + 2.0 * d["dfma"]) / args.npts
        counters[name] = {"thread_inst": d, "points": args.npts, "flops_per_eval": flops[c]}
    js = {"flops_per_eval": flops, "order": [KID[k][0] for k in (7, 8, 9, 18, 19, 20, 21)],
          "definition": "ALGORITHMIC flops of one QAGS integrand evaluation K(x,y,xi) - K(y,x,xi): dadd + dmul + 2 dfma thread "
                        "instructions of tkb_eval_integrand (straight evaluation of the reference formulas, four half kernels "
                        "in full, -fmad=false) per point, measured with ncu on %d points (tools/calibrate_k2_flops.py)" % args.npts,
          "source": args.calibration_source or args.calibration, "counters": counters}
    if args.executed_json:
        ex = json.load(open(args.executed_json))
        js["executed_flops_per_eval"] = [0.0] * 7
        js["fp64_pipe_active_ncu_pct"] = [0.0] * 7
        for k in ex["kernels"]:
            if k.get("kid") in KID:
                for q in (k.get("pair") or [k["kid"]]):          # a paired launch reports for both of its constants
                    c = KID[q][1]
                    js["executed_flops_per_eval"][c] = k.get("executed_flops_per_eval")
                    js["fp64_pipe_active_ncu_pct"][c] = k["fp64_pipe_active_pct"]
        js["fp64_pipe_active_ncu_time_weighted_pct"] = ex.get("k2_fp64_pipe_active_pct_time_weighted")
        js["executed_source"] = args.executed_json
        js["traffic_bytes"] = max(k["dram_bytes"] for k in ex["kernels"] if k.get("kid") in KID)
        js["traffic_source"] = "largest dram__bytes_read.sum + dram__bytes_write.sum of the K2 launches in %s: the kernels have no HBM traffic to speak of (interval-list spill and results only)" % args.executed_json
    with open(args.out, "w") as f:
        json.dump(js, f, indent=1)
    print(json.dumps(js, indent=1))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--raw")
    ap.add_argument("--evals")
    ap.add_argument("--note")
    ap.add_argument("--calibration")
    ap.add_argument("--calibration-source")
    ap.add_argument("--npts", type=int, default=1 << 20)
    ap.add_argument("--executed-json")
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    if a.calibration:
        calibrate(a)
    else:
        summarize(a)
