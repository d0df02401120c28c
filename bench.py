#!/usr/bin/env python3
"""bench.py -- GCUPS of the NW/SW fill+traceback hot path (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--pairs P]

Workload (config.workload): BASELINE.json configs[1] -- a synthetic batch of P (default 1 M) pairs,
150 bp pattern x 1 kb text DNA, seed 481 (+rank), scoring 1/-1/-1, global AND local, score +
traceback.  One "step" = one pass of the hot path over the batch in both modes.  For N > 1 every
rank holds its own P-pair shard (pair-sharded, no data-path collective, "weak" scaling).

  value      = cells / device time of the fill+traceback kernels, inputs already resident in HBM
               (CUDA events recorded by the library on its launching stream), max over ranks.
  e2e        = same metric through b2a_align_batch with HOST (pinned) buffers: H2D of the
               sequences, both kernels, D2H of the result records, every step.
  roofline   = integer-ALU roofline of the dominant kernel (short16 fill): algorithmic int16
               lane-ops (5 per NW cell, 6 per SW cell) / live fill time, against the
               s16x2 DPX issue rate microbenchmarked on this very GPU in this very run.
  cpu_baseline / --impl reference = the UNMODIFIED reference binary (oracle/_ref/hw2) on a bounded
               sample of the same batch on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

M, N_TXT = 150, 1000
SCORING = (1, -1, -1)
OPS_PER_CELL = {0: 5, 1: 6}          # NW 5, SW 6 integer ops per cell


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def bind_to_gpu_numa_node(index):
    """Best effort: run this rank (and allocate its pinned host buffers) on the CPUs next to its GPU.  With 8 ranks the
    end-to-end arm is bounded by the host side of the copies; buffers on the far socket halve the per-GPU H2D rate."""
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if bus.startswith("0000"):
            bus = bus[4:]                                  # sysfs uses a 4-digit PCI domain
        cpus = open(f"/sys/bus/pci/devices/{bus}/local_cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            lo, _, hi = part.partition("-")
            ids.update(range(int(lo), int(hi or lo) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
            return cpus
    except Exception:
        pass
    return None


def reference_run(pairs_per_proc, seed):
    """Times oracle/_ref/hw2 (the unmodified reference) on a bounded sample: one process per host core, each on its
    own pairs_per_proc-pair shard of the same synthetic workload, -g then -l.  Returns aggregate GCUPS + details."""
    import oracle_binding as ob
    from __graft_entry__ import load_package
    load_package()
    from bioinformatics_algorithms_b200 import workload
    if not ob.have_ref():
        return None
    cores = os.cpu_count() or 1
    out = {}
    with tempfile.TemporaryDirectory() as td:
        for c in range(cores):
            pat, po, txt, to = workload.config2(pairs_per_proc, seed=seed + 1000 + c)
            workload.write_fasta(os.path.join(td, f"p{c}.fa"), pat, po, b"p")
            workload.write_fasta(os.path.join(td, f"t{c}.fa"), txt, to, b"t")
        cells = cores * pairs_per_proc * M * N_TXT
        total = 0.0
        for flag in ("-g", "-l"):
            t0 = time.perf_counter()
            procs = [subprocess.Popen([ob.REF_HW2, flag, "-p", os.path.join(td, f"p{c}.fa"), "-t", os.path.join(td, f"t{c}.fa"),
                                       "-o", os.path.join(td, f"o{c}.txt"), "-s", *map(str, SCORING)]) for c in range(cores)]
            for p in procs:
                if p.wait() != 0:
                    raise RuntimeError("reference hw2 failed")
            dt = time.perf_counter() - t0
            out[flag] = cells / dt / 1e9
            total += dt
    return {"value": 2 * cells / total / 1e9, "unit": "GCUPS", "cores": cores, "kind": "reference",
            "sample": f"{cores} processes x {pairs_per_proc} pairs 150x1000 each, -g then -l, oracle/_ref/hw2 (g++ -O2, 1 thread/process)",
            "gcups_global": out["-g"], "gcups_local": out["-l"], "seconds": total}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--pairs", type=int, default=1_000_000, help="pairs per GPU")
    ap.add_argument("--ref-pairs", type=int, default=1500, help="reference sample: pairs per host process and mode")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scoring", default="1,-1,-1", help="match,mismatch,gap (SURVEY 8d also names 2,-3,-4: 4-bit deltas, twice the record)")
    args = ap.parse_args()

    global SCORING
    SCORING = tuple(int(x) for x in args.scoring.split(","))
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    workload_name = f"config2: {args.pairs} pairs/GPU, 150 bp x 1 kb DNA, seed 481+rank, -s {' '.join(map(str, SCORING))}, global+local, score+traceback"

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = max(1, args.steps)
        vals = [reference_run(max(200, args.ref_pairs // 2), 481 + s) for s in range(args.warmup and 1)]  # one short warm-up pass
        vals = [reference_run(args.ref_pairs, 481 + 10 * s) for s in range(steps)]
        if vals[0] is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/hw2 was not built (reference sources absent at build time)"}))
            return 0
        v = float(np.mean([x["value"] for x in vals]))
        secs = float(np.mean([x["seconds"] for x in vals]))
        cb = dict(vals[-1]); cb["value"] = v
        print(json.dumps({"impl": "reference", "metric": "GCUPS (cell updates/s) NW/SW fill+traceback", "value": v, "unit": "GCUPS",
                          "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": secs * 1e3,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                          "config": {"workload": workload_name, "reference_sample": cb["sample"]},
                          "cpu_baseline": cb,
                          "e2e": {"value": v, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    pkg = load_package()
    from bioinformatics_algorithms_b200 import workload

    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout at the first communicator; stdout must carry the JSON line only
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from bioinformatics_algorithms_b200.sharding import max_over_ranks, sum_over_ranks

    n_pairs = args.pairs
    pat_np, po_np, txt_np, to_np = workload.config2(n_pairs, seed=481 + rank)
    # pinned host copies: the e2e path copies from these every step
    pat = pkg.pinned_empty(len(pat_np), np.uint8); pat[:] = pat_np
    txt = pkg.pinned_empty(len(txt_np), np.uint8); txt[:] = txt_np
    po = pkg.pinned_empty(len(po_np), np.uint64); po[:] = po_np
    to = pkg.pinned_empty(len(to_np), np.uint64); to[:] = to_np
    res_host = pkg.pinned_empty(n_pairs, pkg.RESULT_DTYPE)
    del pat_np, txt_np
    cells_mode = n_pairs * M * N_TXT

    eng = {mode: pkg.Engine(local_rank) for mode in (pkg.GLOBAL, pkg.LOCAL)}
    for mode in eng:
        eng[mode].upload(mode, pat, po, txt, to, *SCORING, want_ops=True)

    # ---- device-resident arm: W warm-up steps, then exactly K timed steps ----
    for _ in range(args.warmup):
        for mode in eng:
            eng[mode].run()
    barrier()
    sampler = ClockSampler(local_rank); sampler.start()
    launches0 = sum(eng[mode].stats()["launches"] for mode in eng)
    fill_ms = {0: 0.0, 1: 0.0}; tb_ms = {0: 0.0, 1: 0.0}; tot_ms = {0: 0.0, 1: 0.0}
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for mode in eng:
            eng[mode].run()
            f, t, tot = eng[mode].times()       # CUDA events on the launching streams: per kernel, and first start -> last end
            fill_ms[mode] += f; tb_ms[mode] += t; tot_ms[mode] += tot
    barrier()
    wall_dev = time.perf_counter() - t0
    clocks = sampler.summary()
    dev_s = sum(tot_ms.values()) * 1e-3
    dev_s = max_over_ranks(dev_s)
    wall_dev = max_over_ranks(wall_dev)
    launches = sum(eng[mode].stats()["launches"] for mode in eng) - launches0
    total_cells = sum_over_ranks(2.0 * cells_mode * args.steps)
    value = total_cells / dev_s / 1e9
    results_check = eng[pkg.LOCAL].download(n_pairs)

    # ---- e2e arm: host buffers in, result records out, every step ----
    st0 = {mode: eng[mode].stats() for mode in eng}
    for mode in eng:
        eng[mode].align_packed(mode, pat, po, txt, to, *SCORING, want_ops=True, results=res_host)     # warm-up
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for mode in eng:
            eng[mode].align_packed(mode, pat, po, txt, to, *SCORING, want_ops=True, results=res_host)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = total_cells / e2e_s / 1e9
    st1 = {mode: eng[mode].stats() for mode in eng}
    h2d = sum(st1[m]["h2d_bytes"] for m in eng)         # counters are per last batch
    d2h = sum(st1[m]["d2h_bytes"] for m in eng)
    assert np.array_equal(res_host["score"], results_check["score"]), "e2e and device-resident arms disagree"

    line = None
    if rank == 0:
        # ---- roofline of the dominant kernel (short16 fill), against the DPX issue rate measured live ----
        peak_gops, _ = eng[pkg.GLOBAL].microbench(0)
        mix_gops, _ = eng[pkg.GLOBAL].microbench(1)
        mix2_gops, _ = eng[pkg.GLOBAL].microbench(2)
        peak_cellops = peak_gops * 2.0                   # two int16 cells per 32-bit lane instruction
        alg_ops = sum(cells_mode * args.steps * OPS_PER_CELL[m] for m in eng)
        fill_s = (fill_ms[0] + fill_ms[1]) * 1e-3
        achieved = alg_ops / fill_s / 1e9
        fill_bytes = eng[pkg.GLOBAL].stats()["fill_bytes"]
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        # DRAM traffic of the dominant kernel from the committed ncu --set full capture, scaled per pair to this launch
        traffic, traffic_src = None, None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01_fill_traffic.json")))
            per_pair = sum(tr[m]["dram_read_bytes"] + tr[m]["dram_write_bytes"] for m in ("global", "local")) / 2.0 / tr["pairs_per_launch"]
            traffic, traffic_src = per_pair * n_pairs, tr["source"]
        except Exception:
            pass
        roofline = {"bound": "int_alu", "achieved": achieved, "peak": peak_cellops, "unit": "G int16-cell-ops/s",
                    "frac": achieved / peak_cellops, "traffic": traffic, "traffic_unit": "DRAM bytes per fill launch (mean of NW and SW)",
                    "traffic_source": traffic_src, "algorithmic_bytes_per_launch": fill_bytes,
                    "peak_source": "b2a_microbench_int16x2 kind 0 (VIADDMNMX.S16x2, 8-way ILP, all SMs) measured in this run, x2 cells/lane",
                    "alu_mix_gops": mix_gops, "alu_mix_plus_imad_gops": mix2_gops,
                    "ops_per_cell": {"global": 5, "local": 6},
                    "per_mode": {"global": {"fill_ms": fill_ms[0] / args.steps, "traceback_ms": tb_ms[0] / args.steps,
                                            "gcups_fill": cells_mode * args.steps / (fill_ms[0] * 1e-3) / 1e9,
                                            "frac": cells_mode * args.steps * 5 / (fill_ms[0] * 1e-3) / 1e9 / peak_cellops},
                                 "local": {"fill_ms": fill_ms[1] / args.steps, "traceback_ms": tb_ms[1] / args.steps,
                                           "gcups_fill": cells_mode * args.steps / (fill_ms[1] * 1e-3) / 1e9,
                                           "frac": cells_mode * args.steps * 6 / (fill_ms[1] * 1e-3) / 1e9 / peak_cellops}},
                    "hbm": {"fill_write_bytes_per_mode": fill_bytes,
                            "fill_write_gbs": fill_bytes * 2 * args.steps / fill_s / 1e9,
                            # traceback: DRAM bytes read per pair from the committed ncu capture (profiles/r01_ncu_short16_summary.md)
                            "traceback_read_bytes_per_pair": 7600, "traceback_read_gbs": 7600.0 * n_pairs * 2 * args.steps / ((tb_ms[0] + tb_ms[1]) * 1e-3) / 1e9,
                            "peak_gbs": hbm_peak, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cpu = reference_run(args.ref_pairs, 481)
        line = {"metric": "GCUPS (cell updates/s) NW/SW fill+traceback", "value": value, "unit": "GCUPS", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_s * 1e3 / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16x2", "data": "synthetic",
                "config": {"workload": workload_name, "pairs_per_gpu": n_pairs, "l2": "inputs+record (>50 GB/mode) far larger than L2",
                           "timing": "library CUDA events on the launching streams (first kernel start -> last kernel end per mode), max over ranks",
                           "e2e_pipeline": "b2a_align_batch cuts the batch into segments (16k pairs doubling to 128k); the H2D copy of segment k+1 overlaps the kernels of segment k",
                           "wall_ms_per_step_device_arm": wall_dev * 1e3 / args.steps, "rank0_cpu_binding": numa_cpus},
                "e2e": {"value": e2e_value, "unit": "GCUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_s * 1e3 / args.steps},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
    for e in eng.values():
        e.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
