"""Multi-GPU drivers for configs 3 and 5 (one process per GPU, torchrun; no data-path collective):

  torchrun --nproc-per-node N scripts/run_sharded.py c3 [--pairs 200000]     pair-sharded batch + merged winner (hw2.cpp:326-357)
  torchrun --nproc-per-node N scripts/run_sharded.py c5 [--len 100000]       120-pair all-vs-all distance stage, star sums, centre

Each rank owns a contiguous pair range (sharding.pair_range), runs it on its GPU through the C ABI, and only the
per-rank winner / 16 partial sums cross ranks.  Rank 0 prints one JSON line per config with the max-over-ranks time."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from __graft_entry__ import load_package
pkg = load_package()
from bioinformatics_algorithms_b200 import workload, sharding

ap = argparse.ArgumentParser()
ap.add_argument("what", choices=["c3", "c5"])
ap.add_argument("--pairs", type=int, default=200_000)
ap.add_argument("--len", type=int, default=100_000)
args = ap.parse_args()
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    saved = os.dup(1); os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr)); dist.barrier(); torch.cuda.synchronize()
    sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
eng = pkg.Engine(lr)

if args.what == "c3":
    # every rank regenerates the same global batch and keeps only its range (the CLI does the same from one FASTA pair)
    pat, po, txt, to = workload.config2(args.pairs, seed=481, tie_fraction=0.05)
    first, count = sharding.pair_range(args.pairs, rank, world)
    m, n = 150, 1000
    p_sh, t_sh = pat[first * m:(first + count) * m], txt[first * n:(first + count) * n]
    po_sh = (np.arange(count + 1, dtype=np.uint64) * np.uint64(m)); to_sh = (np.arange(count + 1, dtype=np.uint64) * np.uint64(n))
    out = {}
    for mode, name in ((pkg.GLOBAL, "global"), (pkg.LOCAL, "local")):
        eng.align_packed(mode, p_sh, po_sh, t_sh, to_sh, 1, -1, -1, want_ops=True)     # warm-up
        if world > 1: dist.barrier()
        t0 = time.perf_counter()
        res = eng.align_packed(mode, p_sh, po_sh, t_sh, to_sh, 1, -1, -1, want_ops=True)
        best, key = sharding.merge_best(mode, res, first)
        dt = sharding.max_over_ranks(time.perf_counter() - t0)
        out[name] = {"winner": best, "key": key, "ms": dt * 1e3, "gcups_e2e": args.pairs * m * n / dt / 1e9}
    if rank == 0:
        print(json.dumps({"config": "c3", "n_gpus": world, "pairs": args.pairs, **out}))
else:
    seqs = [x.tobytes() for x in workload.config5(16, args.len, seed=483)]
    first, count = sharding.star_pair_range(len(seqs), rank, world)
    eng.affine_star_scores(seqs[:2], 5, -4, -16, -4)                                     # warm-up (allocations)
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    ps, part, _ = eng.affine_star_scores(seqs, 5, -4, -16, -4, pair_first=first, pair_count=count)
    sums, centre = sharding.reduce_star_sums(part)
    dt = sharding.max_over_ranks(time.perf_counter() - t0)
    cells = sum(len(seqs[i]) * len(seqs[j]) for i in range(16) for j in range(i + 1, 16))
    if rank == 0:
        print(json.dumps({"config": "c5 affine 5:-4:-16:-4", "n_gpus": world, "pairs": 120, "ms": dt * 1e3, "gcups_e2e": cells / dt / 1e9,
                          "centre": int(centre), "sums": [int(x) for x in sums]}))
eng.close()
if world > 1:
    dist.barrier(); dist.destroy_process_group()
