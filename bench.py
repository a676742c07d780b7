#!/usr/bin/env python
"""Benchmark of the hot path: fidelity-kernel Gram entries per second (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[2], SURVEY.md §8d "C3"): FidelityKernel Gram matrix,
ChebyshevPQC(16 qubits, 2 layers), N = 4096 synthetic points (d = 4, seed 3).  One step = simulate
all N statevectors + the N x N Gram.  Prints ONE JSON line (see the keys below).

  value     whole-job entries/s with the features already resident in HBM (K stays in HBM)
  e2e       same metric through FidelityKernel.evaluate(x_host) -> host K: H2D of x and D2H of K
            inside the timed region
  roofline  the Gram kernel against the FP64 GEMM peak measured live on this GPU (cuBLAS DGEMM
            through torch.matmul; MEASURED_PEAKS.json has no FP64 figure), plus `roofline_gate`
            for the gate-application passes against the measured HBM copy bandwidth
  cpu_baseline  the oracle's reference-faithful mode (per-gate einsum + Python pair loop) timed on
            this box's host cores on a bounded sample, extrapolated to the full workload

`--impl reference` times only that CPU arm (the reference itself cannot be imported in this image:
pennylane/qiskit/qulacs are absent) with the same metric/config.
Under torchrun (N > 1) the Gram is sharded by row blocks over the ranks (strong scaling).
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fidelity_kernel_gram_entries_per_sec"
UNIT = "entries/s"


def workload(args):
    n, layers, d, n_points = args.qubits, 2, 4, args.points
    rng = np.random.default_rng(3)
    X = rng.uniform(-0.95, 0.95, (n_points, d))
    return n, layers, d, n_points, X


def config_dict(args, n_gpus):
    return {
        "workload": f"FidelityKernel Gram, ChebyshevPQC({args.qubits} qubits, 2 layers), "
                    f"N={args.points}, d=4 (BASELINE configs[{4 if args.qubits >= 20 else 2}])",
        "n_qubits": args.qubits, "n_points": args.points, "n_gates": 6 * args.qubits,
        "sharding": "single GPU" if n_gpus == 1 else f"row blocks over {n_gpus} ranks; symmetric half-ring of Gram blocks; the 6 int8 nat digit planes of the needed row blocks are pushed peer-to-peer by the copy engines over NVLink (CUDA IPC) and multiplied as they land by one job-table Gram launch per rank (NCCL for the mirror exchange and barriers)",
        "l2_policy": "inputs larger than L2 (statevector batch 16*N*2^n bytes >> 126 MB)",
    }


# ----------------------------------------------------------------------------- CPU arm
def cpu_reference_sample(args, n_states):
    """One bounded sample of the reference algorithm; returns (t_state, t_pair) in seconds."""
    import oracle

    n, layers, d, n_points, X = workload(args)
    gates = oracle.chebyshev_pqc(n, layers, d)
    p = oracle.chebyshev_pqc_initial_parameters(n, layers, d, seed=0)
    t0 = time.perf_counter()
    sv = oracle.simulate_reference_style(gates, n, X[:n_states], p)
    t1 = time.perf_counter()
    oracle.fidelity_gram_pairloop(sv, sv, True)
    t2 = time.perf_counter()
    pairs = n_states * (n_states - 1) / 2
    return (t1 - t0) / n_states, (t2 - t1) / max(pairs, 1)


def cpu_extrapolate(n_points, t_state, t_pair, workers=1):
    """States and pairs are independent units: `workers` processes split them evenly."""
    total = (n_points * t_state + n_points * (n_points - 1) / 2 * t_pair) / max(workers, 1)
    return n_points * n_points / total, total


def cpu_sample_size(args):
    # ~0.16 s per 16-qubit state on one core -> about 8-10 s per sample
    per_state = 0.16 * 2.0 ** (args.qubits - 16)
    return int(max(4, min(args.points, 64, round(9.0 / per_state))))


def cpu_workers():
    """Host processes of the CPU arm: the reference algorithm is single-threaded Python + NumPy per state and
    per pair, so the way to use every host core is one process per core over disjoint states / pairs."""
    return int(max(1, min(os.cpu_count() or 1, int(os.environ.get("B200SQ_CPU_WORKERS", "64")))))


def _cpu_worker(packed):
    qubits, points, n_states = packed

    class _A:
        pass

    a = _A()
    a.qubits, a.points = qubits, points
    return cpu_reference_sample(a, n_states)


def cpu_reference_parallel(args, n_states, workers):
    """`workers` processes run the bounded sample concurrently (same contention for memory bandwidth as a
    full multi-process run would see); returns the mean (t_state, t_pair) and the wall time."""
    if workers <= 1:
        t0 = time.perf_counter()
        ts, tp = cpu_reference_sample(args, n_states)
        return ts, tp, time.perf_counter() - t0
    import multiprocessing as mp

    # one BLAS / OpenMP thread per process (the children inherit the environment at spawn): the processes
    # already cover the cores, nested thread pools only thrash
    saved = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")}
    for k in saved:
        os.environ[k] = "1"
    t0 = time.perf_counter()
    try:
        with mp.get_context("spawn").Pool(workers) as pool:
            res = pool.map(_cpu_worker, [(args.qubits, args.points, n_states)] * workers)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    wall = time.perf_counter() - t0
    return float(np.mean([r[0] for r in res])), float(np.mean([r[1] for r in res])), wall


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_states = cpu_sample_size(args)
    workers = cpu_workers()
    for _ in range(args.warmup):
        cpu_reference_sample(args, max(4, n_states // 8))
    ts, tp, t_steps = [], [], []
    for _ in range(args.steps):
        a, b, wall = cpu_reference_parallel(args, n_states, workers)
        t_steps.append(wall)
        ts.append(a)
        tp.append(b)
    t_state, t_pair = float(np.mean(ts)), float(np.mean(tp))
    value, total = cpu_extrapolate(args.points, t_state, t_pair, workers)
    sample = (f"per step: {workers} processes, each {n_states} states simulated with one einsum per gate on the "
              f"(N,2,..,2) array + the j<i Python pair loop on its {n_states}x{n_states} block "
              f"({t_state:.3f} s/state, {t_pair * 1e6:.0f} us/pair under that load); extrapolated to N={args.points} "
              f"split over the {workers} processes: {total:.0f} s for the full Gram")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(t_steps)), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
        "impl": "reference", "config": config_dict(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample,
                         "host_cores_available": os.cpu_count(), "numpy": np.__version__},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "power_w_max": float(max(power)) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- GPU arm
def fp64_peak_probe(torch, dev):
    """cuBLAS DGEMM 8192^3 through torch.matmul: the FP64 roofline denominator (burst)."""
    m = 8192
    a = torch.randn((m, m), dtype=torch.float64, device=dev)
    b = torch.randn((m, m), dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    torch.cuda.synchronize(dev)
    best = 1e30
    for _ in range(3):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        torch.matmul(a, b)
        e.record()
        torch.cuda.synchronize(dev)
        best = min(best, s.elapsed_time(e))
    del a, b
    return 2.0 * m ** 3 / (best * 1e-3) / 1e12


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    import squlearn_b200 as sq
    from squlearn_b200.distributed import ShardedFidelityKernel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    multi = world > 1
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if multi:
        dist.init_process_group("nccl", device_id=dev)

    n, layers, d, n_points, X = workload(args)
    dim = 1 << n
    executor = sq.Executor("b200", device=local_rank)
    eng = executor.engine
    enc = sq.ChebyshevPQC(n, layers)
    kernel = sq.FidelityKernel(enc, executor)
    kernel._initialize_kernel(d)
    cprog = kernel._circuit.compiled(eng, kernel.parameters)
    n_passes = cprog.num_passes
    sharded = ShardedFidelityKernel(kernel) if multi else None

    x_dev = torch.from_numpy(X).to(dev)
    k_buf = torch.empty((n_points, n_points), dtype=torch.float64, device=dev) if not multi else None
    ev = lambda: torch.cuda.Event(enable_timing=True)
    gate_ms, gram_ms = [], []

    def step_single(record=False):
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        psi = eng.simulate(cprog, x_dev)
        e1.record()
        eng.gram(psi, None, diagonal="one", out=k_buf)
        e2.record()
        if record:
            return (e0, e1, e2)
        return None

    def step_multi(record=False):
        sharded.evaluate_rows(X)
        return None

    step = step_multi if multi else step_single

    def sync_all():
        if multi:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    sync_all()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count()
    recs = []
    t_start, t_end = ev(), ev()
    sync_all()
    t_start.record()
    for _ in range(args.steps):
        recs.append(step(record=True))
    t_end.record()
    sync_all()
    launches = eng.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = t_start.elapsed_time(t_end)
    if multi:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt)
        launches = int(lt.item())
    ms_per_step = elapsed_ms / args.steps
    value = n_points * n_points / (ms_per_step * 1e-3)
    for r in recs:
        if r is not None:
            gate_ms.append(r[0].elapsed_time(r[1]))
            gram_ms.append(r[1].elapsed_time(r[2]))

    # ---- end to end through the public API (host x -> host K)
    e2e_steps = max(2, min(args.steps, 5))

    def e2e_step():
        if multi:
            return sharded.evaluate(X)
        return kernel.evaluate(X)

    k_prev = e2e_step()
    k_prev2 = e2e_step()
    e2e_step()  # warm-ups: the result owns its pinned host buffer, so keep two alive until the pinned
    del k_prev, k_prev2  # allocator has the two blocks the steady-state loop alternates between
    sync_all()
    t0 = time.perf_counter()
    e_start, e_end = ev(), ev()
    e_start.record()
    per_step = []
    for _ in range(e2e_steps):
        ts = time.perf_counter()
        k_host = e2e_step()
        per_step.append(round((time.perf_counter() - ts) * 1e3, 2))
    e_end.record()
    if os.environ.get("B200SQ_BENCH_DEBUG"):
        print(f"rank {rank} e2e per-step wall ms: {per_step}", file=sys.stderr, flush=True)
    sync_all()
    e2e_ms = e_start.elapsed_time(e_end) / e2e_steps
    wall_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    e2e_ms = max(e2e_ms, wall_ms)  # host-side work (D2H wait, numpy) counts too
    if multi:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())

    # ---- parity of what was just timed (outside the timed region): a 256 x 256 block of the host K against the
    # FP64 DMMA Gram of the same states, on every rank count
    parity = None
    if rank == 0:
        nb = min(256, n_points)
        idx = np.arange(0, n_points, max(n_points // nb, 1))[:nb]
        psi_chk = kernel.states(X[idx])
        k64 = eng.gram(psi_chk, None, diagonal="one", precision="fp64").cpu().numpy()
        err = float(np.max(np.abs(np.asarray(k_host)[np.ix_(idx, idx)] - k64)))
        parity = {"max_abs_err_vs_fp64_gram": err, "points": int(nb), "bound": 1e-10,
                  "what": "K[idx, idx] of the end-to-end result vs the FP64 DMMA Gram of the same states"}
        assert err < 1e-10, f"bench result differs from the FP64 Gram: {err}"
        del psi_chk

    line = None
    dtype_str = ("f64 (complex128 states; Gram via exact int8 slices of a 47-bit fixed-point split, FP64 recombination)"
                 if eng.gram_precision(n_points // world if multi else n_points, n_points, dim) == "int8"
                 else "f64 (complex128)")
    if rank == 0:
        assert k_host is not None and k_host.shape == (n_points, n_points)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": dtype_str, "data": "synthetic",
            "config": config_dict(args, world), "clocks": clocks, "gpu_launches": launches, "parity_check": parity,
            "e2e": {"value": n_points * n_points / (e2e_ms * 1e-3), "unit": UNIT,
                    "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(X.nbytes),
                    "d2h_bytes_per_step": int(n_points * n_points * 8)},
        }
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        if not multi:
            fp64_peak = fp64_peak_probe(torch, dev)
            g_ms = float(np.mean(gram_ms))
            traffic = {}
            try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture
                traffic = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json")))
            except Exception:
                pass
            if eng.gram_precision(n_points, n_points, dim) == "int8":
                mn = 256
                tm = (n_points + mn - 1) // mn
                tiles = tm * (tm + 1) // 2
                entries_computed = tiles * mn * mn
                pairs = 22
                ops = 2.0 * pairs * 2 * (2 * dim) * entries_computed  # int8 multiply-adds x 2
                achieved = ops / (g_ms * 1e-3) / 1e12
                bf16 = float(peaks.get("bf16_tflops", 1590.0))
                i8_probe = {}
                try:  # dense tcgen05 kind::i8 ceiling measured on this pool's B200 (tools/i8_peak_probe.cu)
                    i8_probe = json.load(open(os.path.join(ROOT, "profiles", "r2_i8_peak_probe.json")))
                except Exception:
                    pass
                # the kernel is timed inside a long, power-capped step: sustained figure (burst reported beside it)
                i8_peak = float(i8_probe.get("sustained_tops", 2.0 * bf16))
                fp64_equiv = 8.0 * dim * (n_points * (n_points + 1) / 2) / (g_ms * 1e-3) / 1e12
                line["roofline"] = {
                    "kernel": "k_gram_i8<2> (tcgen05.mma kind::i8 on CTA pairs, TMA operands, TMEM s32 accumulators) "
                              "+ k_slice_i8 + k_gram_i8_final", "bound": "tensor",
                    "achieved": achieved, "peak": i8_peak, "unit": "TFLOP/s", "frac": achieved / i8_peak,
                    "traffic": traffic.get("k_gram_i8"), "ms_per_launch": g_ms,
                    "algorithmic": f"split-integer FP64 emulation: {pairs} int8 slice pairs x 2 (Re, Im) x 2*2^n "
                                   f"multiply-adds per entry x {entries_computed} entries computed ({tiles} "
                                   f"lower-triangle 256x256 tiles); the duration is the whole b200sq_gram call "
                                   f"(digit slicing, memset, tcgen05 kernel, |.|^2 kernel)",
                    "peak_source": ("profiles/r2_i8_peak_probe.json: dense tcgen05.mma.cta_group::2.kind::i8 M=N=256 K=32 on "
                                    "all 74 SM pairs, SMEM-resident pseudo-random operands (tools/i8_peak_probe.cu), "
                                    "SUSTAINED over 4 s under the power cap; MEASURED_PEAKS.json has no INT8 figure")
                                   if i8_probe else "2 x MEASURED_PEAKS.json bf16_tflops (no INT8 probe file)",
                    "peak_burst": i8_probe.get("burst_tops"),
                    "frac_of_burst": achieved / float(i8_probe["burst_tops"]) if i8_probe else None,
                    "fp64_equivalent_tflops": fp64_equiv,
                    "fp64_equivalent_note": f"8*2^n flop per entry of the N(N+1)/2 = {n_points * (n_points + 1) // 2} "
                                            f"distinct entries; measured FP64 DGEMM peak here {fp64_peak:.1f} TFLOP/s",
                    "frac_of_fp64_tensor_peak": fp64_equiv / fp64_peak,
                    "share_of_step": g_ms / ms_per_step,
                }
            else:
                mode_4m = os.environ.get("B200SQ_GRAM_MODE", "3m").lower() == "4m"
                bm = 128 if mode_4m else 64
                tiles_m = (n_points + bm - 1) // bm
                # symmetric: only tiles touching the lower triangle run
                tiles = sum(1 for bi in range(tiles_m) for bj in range((n_points + 63) // 64)
                            if bj * 64 <= bi * bm + bm - 1)
                entries_computed = tiles * bm * 64
                flops = 8.0 * dim * entries_computed
                achieved = flops / (g_ms * 1e-3) / 1e12
                line["roofline"] = {
                    "kernel": ("k_gram 4M" if mode_4m else "k_gram3m_ws (3M, warp-specialised)")
                              + ", FP64 DMMA, fused |.|^2 epilogue", "bound": "tensor",
                    "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                    "traffic": traffic.get("k_gram3m_ws"), "ms_per_launch": g_ms,
                    "algorithmic": f"8*2^n flop per entry x {entries_computed} entries computed "
                                   f"({tiles} lower-triangle {bm}x64 tiles of the symmetric Gram)",
                    "executed_tflops": achieved * (1.0 if mode_4m else 0.75),
                    "executed_note": "3M (Gauss) complex product issues 6*2^n tensor flops per entry; the DMMA "
                                     "hardware ceiling measured with tools/dmma_probe.cu is 37.1 TFLOP/s",
                    "peak_source": "measured live: torch.matmul fp64 8192^3 (cuBLAS DGEMM), best of 3; "
                                   "MEASURED_PEAKS.json carries no FP64 figure",
                    "share_of_step": g_ms / ms_per_step,
                }
            hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
            gt_ms = float(np.mean(gate_ms))
            state_bytes = 16.0 * n_points * dim
            layouts = cprog.pass_layouts()
            gate_bytes = sum(16.0 * n_points * ((2.0 ** qi if qi else 0.0) + 2.0 ** qo) for qi, qo in layouts)
            line["roofline_gate"] = {
                "kernel": f"k_tile_pass x {n_passes} (fused gate application, compact active-qubit layouts)",
                "bound": "hbm",
                "achieved": gate_bytes / (gt_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": gate_bytes / (gt_ms * 1e-3) / 1e9 / hbm_peak, "traffic": traffic.get("k_tile_pass"),
                "ms_all_passes": gt_ms, "passes": n_passes,
                "algorithmic": "sum over passes of 16*N*(2^in + 2^out) bytes, (in, out) active qubits per pass = "
                               f"{layouts}: {gate_bytes / 1e9:.2f} GB for {6 * n} gates (the final state is written "
                               "once; qubits still |0> are never stored or read)",
                "full_sweep_equivalent_gbs": state_bytes * (2 * n_passes - 1) / (gt_ms * 1e-3) / 1e9,
                "full_sweep_equivalent_frac": state_bytes * (2 * n_passes - 1) / (gt_ms * 1e-3) / 1e9 / hbm_peak,
                "unfused_equivalent_gbs": 2 * state_bytes * 6 * n / (gt_ms * 1e-3) / 1e9,
                "note": "the last pass is FP64-pipe bound (ncu: profiles/), not HBM bound: it applies 16 one-qubit "
                        "gates and 2 fused phase-table layers per amplitude",
                "fp64_pipe": (lambda ins: {
                    "instructions": ins, "achieved_tinstr_per_s": ins / (gt_ms * 1e-3) / 1e12,
                    "peak_tinstr_per_s": 33.2 / 2.0, "frac": ins / (gt_ms * 1e-3) / 1e12 / (33.2 / 2.0),
                    "peak_source": "measured DFMA rate 33.2 TFLOP/s = 16.6e12 FP64 instructions/s (tools/dmma_probe.cu, "
                                   "profiles/r1_fp64_peak_probe.txt)",
                    "algorithmic": "FP64 instructions of the plan (4 per amplitude and one-qubit gate / phase multiply) "
                                   "summed over the live tiles of every pass",
                })(float(n_points) * sum(cprog.fp64_instructions_per_state())),
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6.65 TB/s",
                "share_of_step": gt_ms / ms_per_step,
            }
            n_states = cpu_sample_size(args)
            workers = cpu_workers()
            t_state, t_pair, _ = cpu_reference_parallel(args, n_states, workers)
            cpu_value, cpu_total = cpu_extrapolate(n_points, t_state, t_pair, workers)
            line["cpu_baseline"] = {
                "value": cpu_value, "unit": UNIT, "cores": workers, "kind": "port",
                "sample": f"{workers} processes, each {n_states} states with one einsum per gate + the Python pair loop "
                          f"on its {n_states}x{n_states} block ({t_state:.3f} s/state, {t_pair * 1e6:.0f} us/pair under "
                          f"that load), extrapolated to N={n_points} over {workers} processes: {cpu_total:.0f} s",
                "host_cores_available": os.cpu_count(),
            }
    if multi:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--qubits", type=int, default=16)
    ap.add_argument("--points", type=int, default=4096)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3  # timing rule: at least three warm-up steps
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
