#!/usr/bin/env python
"""bench.py -- ADVI ELBO-gradient throughput (barcode*timepoint*sample / s) on B200.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  the reference algorithm's CPU path

One "step" = one full ADVI step (K_mc reparameterised draws, log-joint, analytic gradient, fused
optimiser update of all 2D variational parameters) over BASELINE.json configs[1]:
fitness_normal, 10^6 barcodes x 5 time points, 8 MC samples, synthetic counts.  Inputs are
resident in HBM when the timed region starts; the per-step working set (244 MB fp32) exceeds the
126 MB L2, so no explicit flush is needed between iterations.
"""
from __future__ import annotations

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

METRIC = "ADVI ELBO-gradient evals/s (barcode*timepoint*sample/s)"
UNIT = "barcode*timepoint*sample/s"
K_MC = 8
CFG = 2


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self._stop = threading.Event()
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.time(), line.strip()))
                if self._stop.is_set():
                    break
        except Exception:
            pass

    def stop(self):
        self._stop.set()
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self, t0: float, t1: float) -> dict:
        sm, smax, reasons = [], [], set()
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for (_, r) in self.rows[-3:]]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                 f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm)}


def make_workload(world: int, scaling: str):
    """BASELINE configs[1]; weak scaling keeps 10^6 barcodes per GPU (global = world x 10^6)."""
    import barbay_b200 as bb
    spec = dict(bb.synth.CONFIGS[CFG])
    model = spec.pop("model")
    if scaling == "weak" and world > 1:
        spec["n_neutral"] *= world
        spec["n_bc"] *= world
    da, _ = bb.synth.simulate(model, seed=bb.synth.BASE_SEED + CFG, **spec)
    return model, da


def cpu_port_rate(da, budget_s: float, K: int, max_barcodes: int | None = None):
    """Oracle C port (analytic gradient, fp64, OpenMP on all host cores) on a bounded sample of the
    same workload: the first `nb` mutant columns plus all neutrals.  Returns (units/s, cores, sample)."""
    from oracle import cport
    cport.build()
    R = np.asarray(da.bc_count)
    N, M = da.n_neutral, da.n_bc
    nb = M if max_barcodes is None else min(M, max_barcodes)
    T = R.shape[0]

    def run(nb_, steps):
        sub = np.ascontiguousarray(R[:, :N + nb_])
        pp = cport.PortProblem(sub, N, nb_)
        rng = np.random.default_rng(0)
        theta = np.concatenate([rng.standard_normal(pp.D), rng.standard_normal(pp.D)])
        theta[2 * (T - 1) + 2 * nb_:pp.D] += np.log(sub.T.reshape(-1) + 1.0)
        acc = np.full(2 * pp.D, 1e-8)
        pp.advi_steps(theta, acc, 1, K)                     # warm-up (page faults, thread pool)
        t = time.perf_counter()
        pp.advi_steps(theta, acc, steps, K, first_step=1)
        dt = time.perf_counter() - t
        return steps * K * T * (N + nb_) / dt, dt, pp.threads

    probe_nb = min(nb, 100_000)
    rate, dt, cores = run(probe_nb, 1)
    per_step_full = K * T * (N + nb) / rate
    steps = int(max(1, min(50, budget_s / max(per_step_full, 1e-6))))
    if per_step_full > budget_s:                             # shrink the sample instead
        nb = max(probe_nb, int(nb * budget_s / per_step_full))
        steps = 1
    rate, dt, cores = run(nb, steps)
    sample = f"{steps} ADVI step(s) on {N} neutral + {nb} of {M} mutant barcodes x {T} time points, K={K}, fp64, {dt:.1f} s"
    return rate, cores, sample


def run_reference(args, emit):
    """Reference arm (tier rules): the reference algorithm's CPU path on the box's host cores.  Julia is
    absent, so it is the oracle's compiled C/OpenMP port (kind "port") with all host threads.  W warm-up
    and exactly K timed ADVI steps run on a bounded sample of the workload (all neutrals + the first nb
    mutant barcodes) sized so the whole run stays within ~2 minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cport
    cport.build()
    model, da = make_workload(1, "strong")
    R = np.asarray(da.bc_count)
    T, B = R.shape
    N, M = da.n_neutral, da.n_bc
    K = args.mc_samples
    # probe the rate on a small sample, then size the per-step sample
    probe_rate, cores, _ = cpu_port_rate(da, 1.0, K, max_barcodes=50_000)
    budget = 120.0 / max(1, args.steps + args.warmup)
    nb = int(min(M, max(2_000, probe_rate * budget / (K * T) - N)))
    sub = np.ascontiguousarray(R[:, :N + nb])
    pp = cport.PortProblem(sub, N, nb)
    rng = np.random.default_rng(0)
    theta = np.concatenate([rng.standard_normal(pp.D), rng.standard_normal(pp.D)])
    theta[2 * (T - 1) + 2 * nb:pp.D] += np.log(sub.T.reshape(-1) + 1.0)
    acc = np.full(2 * pp.D, 1e-8)
    pp.advi_steps(theta, acc, max(args.warmup, 1), K)
    t0 = time.perf_counter()
    pp.advi_steps(theta, acc, args.steps, K, first_step=args.warmup)
    dt = time.perf_counter() - t0
    units_step = K * T * (N + nb)
    rate = args.steps * units_step / dt
    sample = (f"{args.steps} ADVI steps on {N} neutral + {nb} of {M} mutant barcodes x {T} time points, K={K}, "
              f"fp64, {dt:.1f} s, OpenMP {cores} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"BASELINE configs[1]: fitness_normal, {B} barcodes x {T} time points, {K} MC samples",
                   "optimizer": "DecayedADAGrad",
                   "note": "reference algorithm's CPU path (Julia absent: oracle C port, analytic gradient, OpenMP); "
                           "Turing+ReverseDiff is single-threaded and taped, i.e. slower than this"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--opt", default="decayed", choices=["decayed", "truncated"])
    ap.add_argument("--mc-samples", type=int, default=K_MC)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    # the end-to-end call runs the reference's default number of iterations (ADVI(1, 10_000), src/vi.jl:98)
    ap.add_argument("--e2e-steps", type=int, default=10000)
    args = ap.parse_args()
    # exactly ONE line on stdout: libraries (NCCL prints its version banner there) write to fd 1 too, so fd 1
    # is pointed at stderr for the whole run and the JSON line goes to the saved descriptor
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line: dict):
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    if args.impl == "reference":
        return run_reference(args, emit)

    import torch
    import barbay_b200 as bb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: no CUDA device visible (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    K = args.mc_samples
    model, da = make_workload(world, args.scaling)
    T, B = np.asarray(da.bc_count).shape
    units_step = B * T * K

    eng = bb.Engine(da, model, n_samples=K, dtype=args.dtype, seed=20261018, device=local_rank, rank=rank, world=world)
    if world > 1:
        uid = [bb.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.comm_init(uid[0])
    # a dedicated (non-default) stream: the engine launches on it and the CUDA events below time it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    eng.set_stream(stream.cuda_stream)
    eng.init_params(1)
    if args.opt == "decayed":
        eng.set_optimizer("decayed")
    else:
        eng.set_optimizer("truncated")
    alg_bytes = eng.algorithmic_bytes_per_step

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    eng.step(max(args.warmup, 3))
    barrier()
    launches0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    ev0.record(stream)
    eng.step(args.steps)
    ev1.record(stream)
    barrier()
    t_wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count - launches0
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = units_step * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel, live CUDA-event brackets on the launching stream.
    # For non-hierarchical models the step is two launches: the fused step kernel (pass 2 of step i + pass 1
    # of step i+1: theta, accumulators and counts cross HBM exactly once per step) and the merged tail kernel
    # (partial-sum reduction + shared latents), see DESIGN.md section 4.
    n_prof = min(200, max(10, args.steps))
    ms_tot, ms_p1, ms_p2 = eng.time_steps(n_prof)
    barrier()
    peak, peak_src = measured_peak_gbs()
    t_p2 = ms_p2 / n_prof * 1e-3
    t_p1 = ms_p1 / n_prof * 1e-3
    achieved = alg_bytes / t_p2 / 1e9
    step_achieved = alg_bytes / (ms / args.steps * 1e-3) / 1e9
    fused = t_p1 < 0.1 * t_p2
    roofline = {
        "bound": "hbm",
        "kernel": "pass2_kernel<..., FUSE=1> (gradient + fused optimiser update + next step's pass-1 sums)" if fused
                  else "pass2_kernel (gradient + fused optimiser update)",
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        # dram__bytes_read.sum + dram__bytes_write.sum of that kernel from the committed ncu --set full capture
        # (profiles/r1_final_ncu_fused_kernel.csv; cfg2 fp32 K=8 DecayedADAGrad, 1 GPU) -- null for other configurations
        "traffic": 188.0e6 if (world == 1 and args.dtype == "f32" and args.opt == "decayed" and K == 8) else None,
        "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
        "kernel_us": t_p2 * 1e6, "pass1_us": t_p1 * 1e6, "step_us": ms / args.steps * 1e3,
        "kernel_share_of_step": t_p2 / (ms_tot / n_prof * 1e-3),
        "step_achieved": step_achieved, "step_frac": step_achieved / peak,
        "note": "K=8: issue-bound (two passes regenerate the Philox/Box-Muller noise: 416 warp instructions per "
                "column*sample), DRAM ~15% busy, see profiles/r1_final.md; roofline_k1 is the same measurement at the "
                "reference's default samples_per_step=1",
    }
    # the same measurement at the reference's default samples_per_step = 1 (src/vi.jl:98), for context
    roofline_k1 = None
    if world == 1 and K != 1:
        eng1 = bb.Engine(da, model, n_samples=1, dtype=args.dtype, seed=20261018, device=local_rank)
        eng1.set_stream(stream.cuda_stream)
        eng1.init_params(1)
        eng1.set_optimizer(args.opt)
        eng1.step(20)
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        eng1.step(n_prof)
        a1.record(stream)
        torch.cuda.synchronize()
        step1 = a0.elapsed_time(a1) / n_prof * 1e-3
        _, _, k1_p2 = eng1.time_steps(n_prof)
        ab1 = eng1.algorithmic_bytes_per_step
        roofline_k1 = {"bound": "hbm", "achieved": ab1 / (k1_p2 / n_prof * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                       "frac": ab1 / (k1_p2 / n_prof * 1e-3) / 1e9 / peak, "kernel_us": k1_p2 / n_prof * 1e3,
                       "step_us": step1 * 1e6, "step_achieved": ab1 / step1 / 1e9, "step_frac": ab1 / step1 / 1e9 / peak,
                       "value": B * T / step1, "mc_samples": 1}
        eng1.close()
    clocks = sampler.summary(t_wall0, time.time()) if sampler else None

    # ---- end to end through the public API with HOST buffers: one complete advi()-equivalent call
    # (pack -> bb_create: H2D of counts/maps -> init -> optimiser -> n steps -> ELBO read-back ->
    # bb_get_posterior: D2H), timed on the host clock around the whole call.
    # (all ranks take part; host wall clock, max over ranks)
    n_e2e = args.e2e_steps
    barrier()
    t0 = time.perf_counter()
    eng2 = bb.Engine(da, model, n_samples=K, dtype=args.dtype, seed=20261018, device=local_rank, rank=rank, world=world)
    t_comm = 0.0
    if world > 1:
        tc0 = time.perf_counter()
        uid2 = [bb.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid2, src=0)
        eng2.comm_init(uid2[0])
        t_comm = time.perf_counter() - tc0       # one-time per process: NCCL communicator + CUDA IPC peer mappings
    eng2.init_params(1)
    eng2.set_optimizer(args.opt)
    eng2.step(n_e2e)
    elbo_last = eng2.step(1, elbo_trace=True)            # the step's result read back (8 bytes)
    m, s = eng2.get_posterior()
    dt_e2e = time.perf_counter() - t0
    if dist is not None:
        tt = torch.tensor([dt_e2e, t_comm], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt_e2e, t_comm = float(tt[0].item()), float(tt[1].item())
    n_tot = n_e2e + 1
    h2d = (B * T * 4 + (B * T + 2 * (B - da.n_neutral)) * 4) / world / n_tot     # int32 counts + layout maps per rank
    d2h = (2 * eng2.D * 8) / n_tot + 8.0 / n_tot                                  # posterior (m, sigma) + ELBO
    e2e = {"value": units_step * n_tot / dt_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "steps": n_tot, "seconds": dt_e2e,
           "what": "one complete advi()-equivalent call through the C ABI from HOST arrays on every rank: bb_create "
                   "(pack + H2D of counts / maps) + [comm init] + bb_init_params + bb_set_optimizer + bb_step + "
                   "ELBO read-back + bb_get_posterior (D2H); bytes are amortised over the steps of the call",
           "elbo_last": float(elbo_last[-1]), "posterior_finite": bool(np.isfinite(m).all())}
    if world > 1:
        # context only (the headline `value` above includes it): the communicator set-up is a fixed per-process
        # cost, paid once however many steps (the reference's default max_iters is 10 000) or fits follow
        e2e["comm_init_seconds"] = t_comm
        e2e["value_excluding_comm_init"] = units_step * n_tot / max(dt_e2e - t_comm, 1e-9)
    eng2.close()
    if rank == 0 and world == 1:
        # streaming variant per the base contract: every step re-uploads that step's counts from pinned
        # host memory and reads the step's ELBO back
        cnt_host = torch.from_numpy(np.ascontiguousarray(np.asarray(da.bc_count).astype(np.int32))).pin_memory()
        cnt_dev = torch.empty_like(cnt_host, device="cuda")
        n_s = min(200, args.steps)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_s):
            cnt_dev.copy_(cnt_host, non_blocking=True)
            tr = eng.step(1, elbo_trace=True)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        e2e["streaming"] = {"value": units_step * n_s / (t1 - t0), "unit": UNIT,
                            "h2d_bytes_per_step": int(cnt_host.numel() * 4), "d2h_bytes_per_step": 8 * (K + 1),
                            "steps": n_s, "what": "per step: pinned-host -> device copy of the count matrix + bb_step "
                            "(with ELBO terms) + ELBO read-back"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, cores, sample = cpu_port_rate(da, 15.0, K)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    if sampler:
        sampler.stop()
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling if world > 1 else "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"BASELINE configs[1]: fitness_normal, {B} barcodes x {T} time points, {K} MC samples"
                       + (f" ({B // world} barcodes per GPU)" if world > 1 else ""),
                       "optimizer": "DecayedADAGrad" if args.opt == "decayed" else "TruncatedADAGrad(n=100)",
                       "l2": "inputs_exceed_l2 (per-step working set > 126 MB, no flush needed)",
                       "parallelism": f"barcode-sharded x{world}, one NCCL all-reduce of {5 * T * K} doubles per step"
                       if world > 1 else "single GPU"},
            "roofline": roofline, "roofline_k1": roofline_k1, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        emit(line)
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
