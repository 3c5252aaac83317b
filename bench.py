#!/usr/bin/env python
"""bench.py -- headline measurement of the hot path on B200.

One "step" = one pass of the hot path over one synthetic trace matrix:
    coset LDE (log_blowup 1, shift 31, bit-reversed rows) of a 2^log_rows x width BabyBear trace
    + Poseidon2 MerkleTreeMmcs commit of the resulting 2^(log_rows+1) x width LDE matrix
(BASELINE.json configs[1] and configs[2]: by default 2^23 x 256 -> the 2^24 x 256 LDE matrix is committed).
metric = algorithmic GB/s = (B_lde + B_commit) / t,  B_lde = 4*N*W*(1+2^b),  B_commit = 4*M*W + 32*(2M-1)
(SURVEY.md section 8(d)).  `value`: inputs resident in HBM; `e2e`: the same through the C-ABI with HOST buffers
(pinned trace H2D + root D2H inside the timed region).  With N GPUs every rank proves its own independent
segment-shaped workload (weak scaling, no data-path collective); roots are all-gathered over NCCL.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference      # the CPU restatement (oracle port) on the host cores
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "babybear_lde_plus_poseidon2_commit_throughput"
UNIT = "GB/s"


def algorithmic_bytes(log_rows, width, added_bits):
    n = 1 << log_rows
    m = n << added_bits
    b_lde = 4 * n * width * (1 + (1 << added_bits))
    b_commit = 4 * m * width + 32 * (2 * m - 1)
    return b_lde, b_commit


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region"""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def cpu_sample(log_rows, width, added_bits, steps=1, native=True):
    """oracle port (plain C + OpenMP) on the host cores: LDE + commit of a bounded sample"""
    import numpy as np
    # torchrun exports OMP_NUM_THREADS=1 for every rank; the CPU arm must use all the host threads it can
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    from oracle import oracle as O
    try:
        if native:
            O.build(native=True)
            O.use_native(True)
    except Exception:
        pass
    cores = os.cpu_count() or 1
    n = 1 << log_rows
    trace = O.fill(n * width, 0xB2000000 + (log_rows << 16) + width).reshape(n, width)
    shift = int(O.to_monty([31])[0])
    best = None
    for _ in range(steps):
        t0 = time.perf_counter()
        lde = O.coset_lde_batch(trace, added_bits, shift, bitrev_out=True)
        root, _ = O.merkle_commit([lde])
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    b_lde, b_commit = algorithmic_bytes(log_rows, width, added_bits)
    return (b_lde + b_commit) / best / 1e9, best, cores, root


def run_reference(args):
    """--impl reference: the reference's CPU path.  The real Plonky3/OpenVM prover is Rust with un-vendored crates
    (no cargo here), so this times the oracle port of the same algorithms with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    lr = min(args.log_rows, args.cpu_log_rows)
    vals = []
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_sample(lr, args.width, args.added_bits)
    t_all = time.perf_counter()
    for _ in range(args.steps):
        v, dt, cores, _ = cpu_sample(lr, args.width, args.added_bits)
        vals.append((v, dt))
    wall = time.perf_counter() - t_all
    v = sum(x[0] for x in vals) / len(vals)
    ms = 1e3 * sum(x[1] for x in vals) / len(vals)
    sample = f"coset LDE + Poseidon2 commit of a 2^{lr} x {args.width} trace (log_blowup {args.added_bits}); same per-element work as the 2^{args.log_rows} workload"
    line = {
        "impl": "reference", "metric": METRIC, "value": round(v, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 (BabyBear Montgomery)",
        "data": "synthetic", "config": workload_config(args),
        "cpu_baseline": {"value": round(v, 4), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(v, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": round(wall, 2),
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {"workload": f"coset LDE 2^{args.log_rows} x {args.width} (log_blowup {args.added_bits}, shift 31, bit-reversed) + Poseidon2 MerkleTreeMmcs commit of the 2^{args.log_rows + args.added_bits} x {args.width} LDE matrix",
            "log_rows": args.log_rows, "width": args.width, "log_blowup": args.added_bits,
            "l2": "inputs (>= 8 GB per step) far exceed the 126 MB L2; no explicit flush needed",
            "parallelism": f"{args.gpus} independent segment(s), one per GPU"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log-rows", type=int, default=23)
    ap.add_argument("--width", type=int, default=256)
    ap.add_argument("--added-bits", type=int, default=1)
    ap.add_argument("--cpu-log-rows", type=int, default=21, help="bounded sample size for the CPU baseline")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist

    import zkvm_prover_b200 as z

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = z.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local)
    lib = ctx.lib
    import ctypes as C

    n, w, b = 1 << args.log_rows, args.width, args.added_bits
    shift = z.GENERATOR_MONTY
    seed = 0xB2000000 + (args.log_rows << 16) + w + rank
    trace = ctx.alloc(n, w).fill(seed)
    lde = ctx.alloc(n << b, w)
    dft = z.B200Dft(ctx)
    ctx.sync()
    root = np.zeros(8, np.uint32)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step_resident(events=None):
        if events:
            events[0].record(stream)
        dft.coset_lde_batch(trace, b, shift, bit_reversed=True, out=lde)
        if events:
            events[1].record(stream)
        arr = (C.c_void_p * 1)(lde.h)
        t = C.c_void_p()
        ctx.check(lib.b200zk_merkle_commit(ctx.h, arr, 1, 0, root.ctypes.data, C.byref(t)))  # syncs for the root D2H
        if events:
            events[2].record(stream)
        lib.b200zk_tree_free(ctx.h, t)

    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ctx.launches
    e0, e1 = ev(), ev()
    per_step = []
    e0.record(stream)
    for _ in range(args.steps):
        evs = [ev(), ev(), ev()]
        step_resident(evs)
        per_step.append(evs)
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    launches = ctx.launches - l0
    ms = e0.elapsed_time(e1) / args.steps
    lde_ms = sum(e[0].elapsed_time(e[1]) for e in per_step) / args.steps
    commit_ms = sum(e[1].elapsed_time(e[2]) for e in per_step) / args.steps
    t_max = torch.tensor([ms], device="cuda")
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
        roots = [torch.zeros(8, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(roots, torch.from_numpy(root.astype(np.int64)).cuda())  # gather the Merkle roots of all segments
    ms_max = float(t_max.item())
    b_lde, b_commit = algorithmic_bytes(args.log_rows, w, b)
    value = world * (b_lde + b_commit) / (ms_max * 1e-3) / 1e9

    # ---- dominant kernel: Poseidon2 leaf hashing of the LDE matrix, timed alone with CUDA events
    m_rows = n << b
    d_dig = C.c_void_p()
    ctx.check(lib.b200zk_dev_alloc(ctx.h, m_rows * 32, C.byref(d_dig)))
    for _ in range(2):
        ctx.check(lib.b200zk_hash_rows_dev(ctx.h, lde.h, d_dig))
    ctx.sync()
    k0, k1 = ev(), ev()
    reps = 3
    k0.record(stream)
    for _ in range(reps):
        ctx.check(lib.b200zk_hash_rows_dev(ctx.h, lde.h, d_dig))
    k1.record(stream)
    ctx.sync()
    torch.cuda.synchronize()
    leaf_ms = k0.elapsed_time(k1) / reps
    lib.b200zk_dev_free(ctx.h, d_dig)
    leaf_bytes = 4 * m_rows * w + 32 * m_rows
    peak, peak_src = peaks()
    perms = m_rows * ((w + 7) // 8)
    traffic = None
    try:  # DRAM bytes per launch of this kernel from the committed `ncu --set full` capture of this same workload
        tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        if args.log_rows == 23 and w == 256 and b == 1:
            traffic = tj["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"kernel": "mk::leaf_hash_fast_kernel (Poseidon2 sponge over the LDE rows)", "bound": "hbm", "achieved": round(leaf_bytes / (leaf_ms * 1e-3) / 1e9, 2),
                "peak": peak, "unit": "GB/s", "frac": round(leaf_bytes / (leaf_ms * 1e-3) / 1e9 / peak, 4), "traffic": traffic, "algorithmic_bytes": leaf_bytes, "peak_source": peak_src,
                "ms": round(leaf_ms, 3), "share_of_step": round(leaf_ms / ms, 3), "gperm_per_s": round(perms / (leaf_ms * 1e-3) / 1e9, 3),
                "note": "bound by the integer pipes, not HBM: 564 Montgomery products x 10 FMA-pipe clocks per permutation (32 B absorbed) put the floor at ~6.6 Gperm/s; DESIGN.md section 4",
                "int_pipe": {"gperm_per_s": round(perms / (leaf_ms * 1e-3) / 1e9, 3), "multiply_bound_gperm_per_s": 6.6, "frac": round(perms / (leaf_ms * 1e-3) / 1e9 / 6.6, 3)},
                "lde": {"ms": round(lde_ms, 3), "achieved": round(b_lde / (lde_ms * 1e-3) / 1e9, 2), "frac": round(b_lde / (lde_ms * 1e-3) / 1e9 / peak, 4)},
                "commit_ms": round(commit_ms, 3)}

    # ---- e2e: host buffers through the C ABI (pinned trace -> H2D -> LDE -> commit -> root D2H)
    e2e = None
    if not args.no_e2e:
        host = torch.empty((n, w), dtype=torch.int32).pin_memory()
        ctx.check(lib.b200zk_mat_download(ctx.h, trace.h, host.data_ptr()))
        def step_e2e():
            # the C-ABI call a host-side prover makes: host trace in, root out; the library pipelines the PCIe
            # transfer with the LDE + leaf hashing in column strips (b200zk_lde_commit_host)
            t = C.c_void_p()
            ctx.check(lib.b200zk_lde_commit_host(ctx.h, host.data_ptr(), n, w, b, shift, int(os.environ.get("B200ZK_STRIP", "0")), root.ctypes.data, C.byref(t)))
            lib.b200zk_tree_free(ctx.h, t)

        step_e2e()
        barrier()
        k0, k1 = ev(), ev()
        k0.record(stream)
        for _ in range(args.steps):
            step_e2e()
        k1.record(stream)
        barrier()
        e_ms = torch.tensor([k0.elapsed_time(k1) / args.steps], device="cuda")
        if world > 1:
            dist.all_reduce(e_ms, op=dist.ReduceOp.MAX)
        e2e = {"value": round(world * (b_lde + b_commit) / (float(e_ms.item()) * 1e-3) / 1e9, 3), "unit": UNIT, "h2d_bytes_per_step": 4 * n * w, "d2h_bytes_per_step": 32,
               "ms_per_step": round(float(e_ms.item()), 3)}
        del host

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        lr = min(args.log_rows, args.cpu_log_rows)
        v, dt, cores, _ = cpu_sample(lr, w, b)
        cpu = {"value": round(v, 4), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"oracle port (C+OpenMP restatement, not the p3 binary): LDE + commit of 2^{lr} x {w}, {dt:.2f} s"}

    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(ms_max, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u32 (BabyBear Montgomery)", "data": "synthetic", "config": workload_config(args), "clocks": clocks,
                "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "root": [int(x) for x in root]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
