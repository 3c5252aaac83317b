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


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def cpu_sample(log_rows, width, added_bits, steps=1, native=True):
    """the CPU implementation of the path on the host cores: LDE + commit of a bounded sample.  oracle/bb_oracle.c, built
    -march=native on this box; on an AVX-512 host that is the packed path (16-lane Montgomery arithmetic, Poseidon2 over 16
    rows per permutation, butterflies vectorised along the row -- the structure of Plonky3's CPU code), else the scalar code."""
    import numpy as np
    # torchrun exports OMP_NUM_THREADS=1 for every rank; the CPU arm must use all the host threads it can
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    from oracle import oracle as O
    kind = "scalar C + OpenMP (x86-64-v3 build)"
    try:
        if native:
            O.build(native=True)
            O.use_native(True)
            kind = "AVX-512 packed C + OpenMP (-march=native)" if O.fast_available() else "scalar C + OpenMP (-march=native)"
    except Exception:
        pass
    cores = os.cpu_count() or 1
    n = 1 << log_rows
    trace = O.fill(n * width, 0xB2000000 + (log_rows << 16) + width).reshape(n, width)
    shift = int(O.to_monty([31])[0])
    best = None
    for _ in range(steps):
        t0 = time.perf_counter()
        lde = O.fast_coset_lde_batch(trace, added_bits, shift)
        root, _ = O.fast_merkle_commit_single(lde)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    b_lde, b_commit = algorithmic_bytes(log_rows, width, added_bits)
    return (b_lde + b_commit) / best / 1e9, best, cores, root, kind


def pick_cpu_log_rows(args, budget_s):
    """largest sample height (<= the workload's) whose (warmup + steps) CPU steps fit in `budget_s`, from a 2^17-row probe"""
    probe = 17 if args.log_rows > 17 else args.log_rows
    cpu_sample(probe, args.width, args.added_bits)                      # builds the library, warms the thread pool
    _, dt, _, _, _ = cpu_sample(probe, args.width, args.added_bits)
    per_step = budget_s / max(1, args.steps + min(args.warmup, 1))
    lr = probe
    while lr < args.log_rows and dt * 2.2 <= per_step:                  # ~2.1x per doubling (n log n + linear hashing)
        lr += 1
        dt *= 2.1
    return lr


def run_reference(args):
    """--impl reference: the reference's CPU path.  The real Plonky3/OpenVM prover is Rust with un-vendored crates
    (no cargo here), so this times the oracle port of the same algorithms with all host threads, on a bounded sample of
    the workload: same width, blow-up and shift, as many rows as fit the time budget (stated in config.sample and
    cpu_baseline.sample; GB/s normalises the height, the per-element work differs by the log-height ratio of the NTT)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    lr = min(args.log_rows, args.cpu_log_rows) if args.cpu_log_rows else pick_cpu_log_rows(args, args.cpu_budget_s)
    vals = []
    for _ in range(min(args.warmup, 1)):
        cpu_sample(lr, args.width, args.added_bits)
    t_all = time.perf_counter()
    for _ in range(args.steps):
        v, dt, cores, _, kind = cpu_sample(lr, args.width, args.added_bits)
        vals.append((v, dt))
    wall = time.perf_counter() - t_all
    v = sum(x[0] for x in vals) / len(vals)
    ms = 1e3 * sum(x[1] for x in vals) / len(vals)
    sample = (f"coset LDE + Poseidon2 commit of a 2^{lr} x {args.width} trace (log_blowup {args.added_bits}) per step -- "
              f"{'the full workload' if lr == args.log_rows else f'a bounded sample of the 2^{args.log_rows}-row workload (same width / blow-up / shift)'}; "
              f"{kind}; restated oracle, not the p3 binary; {cores} threads on {cpu_model()}")
    cfg = workload_config(args)
    cfg["sample"] = {"log_rows": lr, "width": args.width, "log_blowup": args.added_bits, "full_workload": lr == args.log_rows}
    line = {
        "impl": "reference", "metric": METRIC, "value": round(v, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 (BabyBear Montgomery)",
        "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": round(v, 4), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "cpu_model": cpu_model()},
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
    ap.add_argument("--cpu-log-rows", type=int, default=0, help="sample height of the CPU arm (0: the largest that fits --cpu-budget-s)")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0, help="time budget of the whole --impl reference run")
    ap.add_argument("--no-grid", action="store_true", help="skip the BASELINE LDE grid")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the column-sharded single-matrix record")
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
    ctx_sms = torch.cuda.get_device_properties(local).multi_processor_count
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
    traffic, traffic_src, lde_traffic = None, None, None
    try:  # DRAM bytes per launch of this kernel: one `ncu --set full` capture of this same command (profiles/, per round)
        tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        if args.log_rows == 23 and w == 256 and b == 1:
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj.get("source", "profiles/roofline_traffic.json (ncu --set full capture of this workload)")
            lde_traffic = tj.get("lde_dram_bytes_per_step")
    except Exception:
        pass
    gperm = perms / (leaf_ms * 1e-3) / 1e9
    # integer roofline of the sponge: 564 Montgomery products per permutation, each 10 clocks of the
    # FMA-heavy pipe per warp (IMAD.WIDE 4 + IMAD 2 + IMAD.HI 4, measured: tools/pipe_microbench*.cu), 4 sub-partitions per SM
    sm_clock_ghz = (clocks.get("sm_mhz") or 1965) / 1e3
    int_peak = ctx_sms * 4 * 32 * sm_clock_ghz / (564 * 10)
    roofline = {"kernel": "mk::leaf_hash_fast_kernel (Poseidon2 sponge over the LDE rows)", "bound": "int32",
                "achieved": round(leaf_bytes / (leaf_ms * 1e-3) / 1e9, 2), "peak": peak, "unit": "GB/s",
                "frac": round(leaf_bytes / (leaf_ms * 1e-3) / 1e9 / peak, 4), "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes": leaf_bytes, "peak_source": peak_src,
                "ms": round(leaf_ms, 3), "share_of_step": round(leaf_ms / ms, 3),
                "achieved_int": round(gperm, 3), "peak_int": round(int_peak, 3), "unit_int": "Gperm/s", "frac_int": round(gperm / int_peak, 3),
                "peak_int_source": "564 products x 10 FMA-heavy-pipe clocks per warp-permutation at the SM clock sampled during the run (pipe rates measured in profiles/pipe_microbench_r01.txt)",
                "note": "achieved/peak/frac are the HBM view the bench contract asks for; the kernel is bound by the 32-bit integer pipes (frac_int), its DRAM traffic equals its algorithmic bytes",
                "lde": {"ms": round(lde_ms, 3), "achieved": round(b_lde / (lde_ms * 1e-3) / 1e9, 2), "frac": round(b_lde / (lde_ms * 1e-3) / 1e9 / peak, 4),
                        "bound": "hbm+int32", "traffic": lde_traffic, "dram_frac": (round(lde_traffic / (lde_ms * 1e-3) / 1e9 / peak, 4) if lde_traffic else None),
                        "note": "traffic = DRAM bytes of the 7 LDE launches of a step (ncu, profiles/ncu_summary_r02.json), dram_frac = traffic / time / peak; 6 pass sweeps of 8 B per element + the fused middle (4 B read + 8 B written per element); per-pass times in profiles/ntt_fused_mid_r02.txt, copy-only ceilings of the tile shapes in profiles/tile_copy_lab_r02.txt"},
                "commit_ms": round(commit_ms, 3)}

    # ---- e2e: host buffers through the C ABI (pinned trace -> H2D -> LDE -> commit -> root D2H)
    e2e = None
    if not args.no_e2e:
        host = torch.empty((n, w), dtype=torch.int32).pin_memory()
        ctx.check(lib.b200zk_mat_download(ctx.h, trace.h, host.data_ptr()))
        # strip width of the pipeline: 0 = the library's choice (32 columns for a blocking call, 64 for the asynchronous stream);
        # B200ZK_STRIP=<cols> forces one (experiments)
        strip_cols = int(os.environ.get("B200ZK_STRIP", "0"))

        def step_e2e():
            # the C-ABI call a host-side prover makes: host trace in, root out; the library pipelines the PCIe
            # transfer with the LDE + leaf hashing in column strips (b200zk_lde_commit_host)
            t = C.c_void_p()
            ctx.check(lib.b200zk_lde_commit_host(ctx.h, host.data_ptr(), n, w, b, shift, strip_cols, root.ctypes.data, C.byref(t)))
            lib.b200zk_tree_free(ctx.h, t)

        def run_serial(k):
            for _ in range(k):
                step_e2e()

        def run_stream(k):
            # the same K commits as a stream, the way a prover walks the segments of a chunk proof: commit i + 1 is issued
            # (b200zk_lde_commit_host_async) before the root of commit i is read back, so the first strip's transfer of one
            # commit runs under the arithmetic of the previous one.  Every commit still moves its whole trace host -> device
            # and its root device -> host inside the timed region; at most two are in flight.
            pending = None
            for _ in range(k):
                t = C.c_void_p()
                ctx.check(lib.b200zk_lde_commit_host_async(ctx.h, host.data_ptr(), n, w, b, shift, strip_cols, C.byref(t)))
                if pending is not None:
                    ctx.check(lib.b200zk_tree_root(ctx.h, pending, root.ctypes.data))
                    lib.b200zk_tree_free(ctx.h, pending)
                pending = t
            ctx.check(lib.b200zk_tree_root(ctx.h, pending, root.ctypes.data))
            lib.b200zk_tree_free(ctx.h, pending)

        def timed(fn, warm=1):
            fn(warm)  # untimed: the allocator reaches its steady state (the stream keeps two LDEs + trees alive at once)
            barrier()
            k0, k1 = ev(), ev()
            k0.record(stream)
            fn(args.steps)
            k1.record(stream)
            barrier()
            t_ms = torch.tensor([k0.elapsed_time(k1) / args.steps], device="cuda")
            if world > 1:
                dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
            return float(t_ms.item())

        root_dev = root.copy()
        serial_ms = timed(run_serial)
        assert np.array_equal(root, root_dev), "host strip pipeline: root differs from the device-resident path"
        stream_note = None
        try:
            stream_ms = timed(run_stream, warm=max(3, args.warmup))
            assert np.array_equal(root, root_dev), "host strip pipeline (async): root differs from the device-resident path"
        except z.B200zkError as ex:  # e.g. no room for two commits in flight: report the blocking number as the e2e value
            ctx.sync()
            stream_ms, stream_note = serial_ms, f"streamed mode failed ({ex}); value is the blocking call's"
        e_ms = torch.tensor([stream_ms], device="cuda")
        # what the host link gives every rank when all ranks copy at once: a plain contiguous pinned H2D copy (2 GiB), all
        # ranks together -- the bound of any e2e number at this N (the GPUs of one box share host memory and PCIe uplinks)
        chunk = host.view(-1)[: min(host.numel(), 1 << 29)]
        dev_chunk = torch.empty_like(chunk, device="cuda")
        dev_chunk.copy_(chunk, non_blocking=True)
        barrier()
        c0, c1 = ev(), ev()
        c0.record()
        for _ in range(2):
            dev_chunk.copy_(chunk, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        link = torch.tensor([2 * chunk.numel() * 4 / (c0.elapsed_time(c1) * 1e-3) / 1e9], device="cuda")
        if world > 1:
            dist.all_reduce(link, op=dist.ReduceOp.MIN)
        del dev_chunk
        link_gbs = float(link.item())
        e2e = {"value": round(world * (b_lde + b_commit) / (float(e_ms.item()) * 1e-3) / 1e9, 3), "unit": UNIT, "h2d_bytes_per_step": 4 * n * w, "d2h_bytes_per_step": 32,
               "ms_per_step": round(float(e_ms.item()), 3), "strip_cols": strip_cols or 64,
               "mode": stream_note or "stream of K commits through b200zk_lde_commit_host_async, at most two in flight; every commit's H2D and root D2H inside the timed region",
               "serial": {"value": round(world * (b_lde + b_commit) / (serial_ms * 1e-3) / 1e9, 3), "ms_per_step": round(serial_ms, 3), "strip_cols": strip_cols or 32,
                          "mode": "one blocking b200zk_lde_commit_host call at a time"},
               "h2d_link_gbs_per_gpu": round(link_gbs, 1), "h2d_floor_ms": round(4 * n * w / link_gbs / 1e6, 1),
               "note": "h2d_link_gbs_per_gpu = slowest rank's plain contiguous pinned copy with all ranks copying at once; h2d_floor_ms = this step's input bytes at that rate"}
        del host

    # ---- the result itself: the root of the default workload must equal the oracle's (computed once on the CPU by
    # tests/golden/make_headline_golden.py).  A fast kernel with a different root is not a result.
    root_check = None
    if rank == 0:
        try:
            g = json.load(open(os.path.join(ROOT, "tests", "golden", f"headline_2p{args.log_rows}x{w}.json")))
            if g["log_blowup"] == b and g["seed"] == seed:
                root_check = {"golden": f"tests/golden/headline_2p{args.log_rows}x{w}.json (oracle, CPU)", "equal": [int(x) for x in root] == g["root"]}
                if not root_check["equal"]:
                    raise SystemExit(f"bench.py: Merkle root {root.tolist()} differs from the oracle's {g['root']} -- refusing to report a throughput")
        except FileNotFoundError:
            pass

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        a1 = argparse.Namespace(**vars(args))
        a1.steps, a1.warmup = 1, 0
        lr = min(args.log_rows, args.cpu_log_rows) if args.cpu_log_rows else pick_cpu_log_rows(a1, 20.0)
        v, dt, cores, _, kind = cpu_sample(lr, w, b)
        cpu = {"value": round(v, 4), "unit": UNIT, "cores": cores, "kind": "port", "cpu_model": cpu_model(),
               "sample": f"LDE + commit of 2^{lr} x {w} (log_blowup {b}) in {dt:.2f} s; {kind}; restated oracle (oracle/bb_oracle.c), not the p3 binary"}

    # ---- BASELINE.json configs[1]: the coset-LDE grid (2^20 - 2^24 rows x 64 - 512 columns), N = 1 only
    grid = None
    if rank == 0 and world == 1 and not args.no_grid:
        grid = []
        trace.free()
        lde.free()
        ctx.trim()
        free_b = torch.cuda.mem_get_info()[0]
        for lr_, w_ in [(20, 64), (20, 256), (20, 512), (22, 64), (22, 256), (22, 512), (24, 64), (24, 256), (24, 512)]:
            need = 4 * (1 << lr_) * w_ * 3
            if need * 1.1 > free_b:
                grid.append({"log_rows": lr_, "width": w_, "skipped": "not enough free device memory"})
                continue
            t_ = ctx.alloc(1 << lr_, w_).fill(0xB2000000 + (lr_ << 16) + w_)
            o_ = ctx.alloc(2 << lr_, w_)
            for _ in range(2):
                dft.coset_lde_batch(t_, 1, shift, bit_reversed=True, out=o_)
            ctx.sync()
            g0, g1 = ev(), ev()
            reps_ = 3
            g0.record(stream)
            for _ in range(reps_):
                dft.coset_lde_batch(t_, 1, shift, bit_reversed=True, out=o_)
            g1.record(stream)
            ctx.sync()
            torch.cuda.synchronize()
            gms = g0.elapsed_time(g1) / reps_
            grid.append({"log_rows": lr_, "width": w_, "ms": round(gms, 3), "GB_per_s": round(need / gms / 1e6, 1), "frac_hbm": round(need / gms / 1e6 / peak, 4),
                         "checksum": f"{o_.checksum():016x}"})
            t_.free()
            o_.free()
        ctx.trim()

    # ---- N > 1: ONE wide matrix, column-sharded (second row): every rank extends W/N columns of
    # the SAME 2^log_rows x W trace rank 0 just proved alone, the last NTT pass stores finished tiles into the row-block
    # owner's memory over NVLink (TMA into a peer mapping), every rank hashes its row block, the N subtree roots are
    # all-gathered (NCCL) and the top log2 N levels finished redundantly.  Strong scaling against this run's own N = 1 step.
    sharded = None
    if world > 1 and not args.no_sharded and w % world == 0 and (world & (world - 1)) == 0:
        from zkvm_prover_b200 import dist as D
        if rank != 0:                       # every rank needs rank 0's matrix: same seed, then its own column slice
            trace.fill(0xB2000000 + (args.log_rows << 16) + w)
        ctx.sync()
        wg = w // world

        class _DevArr:  # zero-copy torch view of the library's matrix (CUDA array interface)
            def __init__(self, ptr, shape):
                self.__cuda_array_interface__ = {"shape": shape, "typestr": "<i4", "data": (ptr, False), "version": 2}

        full = torch.as_tensor(_DevArr(trace.device_ptr, (n, w)), device="cuda")
        view = full[:, rank * wg:(rank + 1) * wg].contiguous()      # this rank's column shard, N x W/G
        torch.cuda.synchronize()
        src = ctx.wrap(view.data_ptr(), n, wg, keepalive=view)
        del full
        trace.free()
        lde.free()
        ctx.trim()
        exch = D.PeerExchange(ctx, n << b, wg)
        times = []
        sroot = None
        for it in range(2 + 3):
            dist.barrier()
            torch.cuda.synchronize()
            s0_, s1_ = ev(), ev()
            s0_.record(stream)
            sroot, _cap = D.sharded_lde_commit_p2p(ctx, src, b, shift, exch)
            s1_.record(stream)
            ctx.sync()
            torch.cuda.synchronize()
            tms = torch.tensor([s0_.elapsed_time(s1_)], device="cuda")
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            if it >= 2:
                times.append(float(tms.item()))
        exch.close()
        s_ms = sum(times) / len(times)
        eq = torch.tensor([1 if [int(x) for x in sroot] == [int(x) for x in roots[0].cpu().tolist()] else 0], device="cuda")
        dist.all_reduce(eq, op=dist.ReduceOp.MIN)
        sharded = {"workload": f"ONE 2^{args.log_rows} x {w} trace, columns sharded over {world} GPUs: LDE + exchange fused into the last NTT pass (TMA stores into peer memory over NVLink) + per-rank subtree + NCCL all-gather of {world} cap digests",
                   "ms": round(s_ms, 3), "speedup_vs_n1": round(ms_max / s_ms, 3), "n1_ms": round(ms_max, 3), "root_equal_single_gpu": bool(eq.item()),
                   "nvlink_bytes": int(4 * (n << b) * w * (world - 1) // world), "GB_per_s": round((b_lde + b_commit) / (s_ms * 1e-3) / 1e9, 1),
                   "scaling": "strong", "all_ms": [round(t, 2) for t in times]}

    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(ms_max, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u32 (BabyBear Montgomery)", "data": "synthetic", "config": workload_config(args), "clocks": clocks,
                "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "root": [int(x) for x in root], "root_check": root_check}
        if grid is not None:
            line["lde_grid"] = grid
        if sharded is not None:
            line["sharded"] = sharded
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
