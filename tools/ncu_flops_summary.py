#!/usr/bin/env python3
"""Turn an ncu counter pass over `bench.py --steps 1 --warmup 3 --no-cpu` into profiles/r02_fp64_flops.json, the file
bench.py reads for its roofline (executed FP64 flop per (E,E',T) triple and per kernel, DRAM bytes, pipe utilisation).

On the GPU box (one call, after the plain run has exited 0):

    ncu --metrics gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,\
smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,\
sm__inst_executed_pipe_tensor_subpipe_dmma.sum,sm__inst_executed_pipe_fp64.sum,\
sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active,\
dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -k regex:"pair_closed|ic_photon|kernels_mma|kernels_wien|solve_pdi" -c 28 --csv --log-file gpurun_out/r02_flops.csv \
        python bench.py --steps 1 --warmup 3 --no-cpu

Here:   python tools/ncu_flops_summary.py gpurun_out/r02_flops.csv

The k-th launch of each kernel belongs to slice k of the benchmark scan (warm-up slices 0-2, timed slice 3); the triples of a
slice (sum over its points of NT * NE(NE+1)/2) are recomputed here from the scan definition.  Flop convention: thread-level
DFMA = 2, DMUL = DADD = 1 (predicated-on), one DMMA.8x8x4 warp instruction = 512."""
import csv
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHORT = ("pair_closed_forms_kernel", "ic_photon_closed_kernel", "kernels_mma_kernel", "kernels_wien_kernel", "solve_pdi_kernel",
         "kernels_shared4_kernel", "direct_rates_kernel")


def source_hash():
    h = hashlib.sha256()
    d = os.path.join(ROOT, "acropolis_b200", "csrc")
    for f in sorted(os.listdir(d)):
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def slice_triples(n):
    import acropolis_b200.flags as flags
    flags.verbose = False
    import bench
    w = bench.WORKLOADS["cfg3"]
    pts = bench.grid_points(w)
    out = []
    for k in range(n):
        mods = [m for m in bench.build_models(w, bench.slice_params(w, pts, k)) if not m.is_trivial()]
        t = 0.0
        for m in mods:
            s = m.point_spec()
            t += len(s.T) * len(s.E_grid) * (len(s.E_grid) + 1) / 2
        out.append(t)
    return out


def main(path, out_path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ix = {h: i for i, h in enumerate(hdr)}
    launches = {}
    for r in rows:
        name = r[ix["Kernel Name"]]
        short = next((s for s in SHORT if s in name), None)
        if short is None:
            continue
        launches.setdefault(short, {}).setdefault(int(r[ix["ID"]]), {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
    # one pair_closed_forms_kernel launch per slice; kernels_mma_kernel / kernels_wien_kernel launch once per integral kind
    n_slices = len(launches.get("pair_closed_forms_kernel") or min(launches.values(), key=len))
    trip = slice_triples(n_slices)
    kernels = {}
    for short, byid in launches.items():
        per = []
        ordered = [m for _, m in sorted(byid.items())]
        per_slice = max(1, len(ordered) // n_slices)       # a kernel launched several times per slice (one launch per integral
        merged = []                                         # kind): its launches of one slice are summed
        for k in range(n_slices):
            grp = ordered[k * per_slice:(k + 1) * per_slice]
            if not grp:
                continue
            tot = {}
            dur = sum(g["gpu__time_duration.sum"] for g in grp)
            for key in grp[0]:
                if key.endswith(".sum"):
                    tot[key] = sum(g.get(key, 0) for g in grp)
                else:                                       # percentages: duration-weighted mean
                    tot[key] = sum(g.get(key, 0) * g["gpu__time_duration.sum"] for g in grp) / dur
            merged.append(tot)
        for k, m in enumerate(merged):
            lid = k
            flop = (2 * m.get("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", 0) + m.get("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", 0) +
                    m.get("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", 0) + 512 * m.get("sm__inst_executed_pipe_tensor_subpipe_dmma.sum", 0))
            per.append(dict(slice=k, triples=trip[k], flop=flop, flop_dmma=512 * m.get("sm__inst_executed_pipe_tensor_subpipe_dmma.sum", 0),
                            ms_under_ncu=m["gpu__time_duration.sum"] / 1e6,
                            fp64_pipe_active_pct=m.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                            shared_pipe_active_pct=m.get("sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active"),
                            dram_bytes=m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0),
                            dram_read=m.get("dram__bytes_read.sum", 0), dram_write=m.get("dram__bytes_write.sum", 0)))
        n = len(per)
        mean = lambda key: sum(p[key] for p in per if p[key] is not None) / n
        kernels[short] = dict(launches=n, launches_per_slice=per_slice, flop_per_triple=sum(p["flop"] for p in per) / sum(p["triples"] for p in per),
                              flop_per_launch=mean("flop"), dmma_share_of_flop=(sum(p["flop_dmma"] for p in per) / max(sum(p["flop"] for p in per), 1)),
                              ms_under_ncu=mean("ms_under_ncu"), tflops_under_ncu=mean("flop") / (mean("ms_under_ncu") * 1e-3) / 1e12,
                              fp64_pipe_active_pct=mean("fp64_pipe_active_pct"), shared_pipe_active_pct=mean("shared_pipe_active_pct"),
                              dram_bytes_per_launch=mean("dram_bytes"), dram_read_per_launch=mean("dram_read"),
                              dram_write_per_launch=mean("dram_write"), per_slice=per)
    res = dict(source=os.path.relpath(path, ROOT), library_source_hash=source_hash(),
               command="bench.py --steps 1 --warmup 3 --no-cpu (cfg3 slices 0-3, 1000 points each)",
               flop_convention="thread-level DFMA x2 + DMUL + DADD (predicated on) + 512 per DMMA.8x8x4 warp instruction",
               note="sm__pipe_shared_cycles_active = the pipe FP64 FMA and FP64 tensor (DMMA) instructions share on B200",
               triples_per_slice=trip, kernels=kernels, when=time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()))
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1)
    for k, v in kernels.items():
        print("%-26s %2d launches  %.3e flop/launch (DMMA share %.2f)  %.2f ms  %.2f TFLOP/s  fp64 pipe %.1f %%  shared pipe %.1f %%  DRAM %.2f GB" % (
            k, v["launches"], v["flop_per_launch"], v["dmma_share_of_flop"], v["ms_under_ncu"], v["tflops_under_ncu"],
            v["fp64_pipe_active_pct"] or 0, v["shared_pipe_active_pct"] or 0, v["dram_bytes_per_launch"] / 1e9))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "r02_fp64_flops.json"))
