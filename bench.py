#!/usr/bin/env python
"""bench.py -- ADVI ELBO-gradient throughput (barcode*timepoint*sample / s) on B200.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  the reference algorithm's CPU path

One "step" = one full ADVI step (K_mc reparameterised draws, log-joint, analytic gradient, fused
optimiser update of all 2D variational parameters) over BASELINE.json configs[1]:
fitness_normal, 10^6 barcodes x 5 time points, 8 MC samples, synthetic counts.  Inputs are
resident in HBM when the timed region starts.

Multi-GPU (torchrun, one rank per GPU): the headline is STRONG scaling -- the same 10^6 barcodes sharded
over the N GPUs (north_star: "10^6 barcodes ... scales >= 6x at 8 GPUs"); the weak-scaling figure
(10^6 barcodes per GPU) is reported inside the line as `scaling_weak`.  Before timing, a sharded run is
checked against an unsharded one on rank 0 (`parity`).
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
SEED = 20261018
L2_BYTES = 126e6


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel_key: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    machine-readable extract of the ncu --set full capture (profiles/ncu_traffic.json: {key: {bytes, source}});
    None when no capture of this build's kernel / configuration is committed."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        ent = json.load(open(p)).get(kernel_key)
        return (float(ent["bytes"]), ent.get("source")) if ent else (None, None)
    except Exception:
        return None, None


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

    def wait_first(self, timeout: float = 8.0):
        """nvidia-smi needs about a second before its first sample: block until it streams (the timed region of the
        default run is 0.3 s long and would otherwise be over before the first row)."""
        t0 = time.time()
        while not self.rows and time.time() - t0 < timeout and self.is_alive():
            time.sleep(0.05)

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


def make_workload(cfg: int, mult: int = 1, scale: float = 1.0, seed_offset: int = 0):
    """BASELINE configs[cfg - 1]; mult > 1 multiplies the barcode axis (weak scaling: 10^6 barcodes per GPU)."""
    import barbay_b200 as bb
    spec = dict(bb.synth.CONFIGS[cfg])
    model = spec.pop("model")
    for key in ("n_neutral", "n_bc"):
        spec[key] = max(8, int(round(spec[key] * mult * scale)))
    if "n_geno" in spec and scale != 1.0:
        spec["n_geno"] = max(2, int(round(spec["n_geno"] * scale)))
    da, _ = bb.synth.simulate(model, seed=bb.synth.BASE_SEED + cfg + seed_offset, **spec)
    return model, da


def count_shape(da):
    R = np.asarray(da.bc_count)
    return R.shape, int(R.size)          # (T, B[, R]), T * B * R


# ------------------------------------------------------------------------------------------ reference arm
def cpu_port_rate(da, budget_s: float, K: int, max_barcodes: int | None = None):
    """Oracle C port (analytic gradient, fp64, OpenMP on all host cores) on a bounded sample of the
    same workload: the first `nb` mutant columns plus all neutrals.  Returns (units/s, cores, sample)."""
    from oracle import cport
    cport.build()
    cores = cport.set_threads()                     # every host core, whatever OMP_NUM_THREADS says (torchrun sets 1)
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
        return steps * K * T * (N + nb_) / dt, dt

    probe_nb = min(nb, 100_000)
    rate, dt = run(probe_nb, 1)
    per_step_full = K * T * (N + nb) / rate
    steps = int(max(1, min(50, budget_s / max(per_step_full, 1e-6))))
    if per_step_full > budget_s:                             # shrink the sample instead
        nb = max(probe_nb, int(nb * budget_s / per_step_full))
        steps = 1
    rate, dt = run(nb, steps)
    sample = (f"{steps} ADVI step(s) on {N} neutral + {nb} of {M} mutant barcodes x {T} time points, K={K}, fp64, "
              f"{dt:.1f} s, OpenMP {cores} threads")
    return rate, cores, sample


def run_reference(args, emit):
    """Reference arm (tier rules): the reference algorithm's CPU path on the box's host cores.  Julia is
    absent, so it is the oracle's compiled C/OpenMP port (kind "port") with ALL host threads (set explicitly:
    torchrun exports OMP_NUM_THREADS=1).  The workload is the repo arm's: BASELINE configs[1], 10^6 barcodes
    (strong scaling: the same problem at every N).  W warm-up and exactly K timed ADVI steps run on a bounded
    sample of it (all neutrals + the first nb mutant barcodes) sized so the whole run stays within ~2 minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cport
    cport.build()
    cores = cport.set_threads()
    mult = args.gpus if args.scaling == "weak" else 1
    model, da = make_workload(CFG, mult)
    R = np.asarray(da.bc_count)
    T, B = R.shape
    N, M = da.n_neutral, da.n_bc
    K = args.mc_samples
    probe_rate, _, _ = cpu_port_rate(da, 1.0, K, max_barcodes=50_000)
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
    rate = args.steps * K * T * (N + nb) / dt
    sample = (f"{args.steps} ADVI steps on {N} neutral + {nb} of {M} mutant barcodes x {T} time points, K={K}, "
              f"fp64, {dt:.1f} s, OpenMP {cores} threads")
    emit({
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
    })


# ------------------------------------------------------------------------------------------ our arm
class Ctx:
    """Process-wide state of one bench run (rank, device, stream, torch.distributed handle)."""

    def __init__(self):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a B200: no CUDA device visible (there is no CPU fallback)")
        torch.cuda.set_device(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
        # a dedicated (non-default) stream: the engine launches on it and the CUDA events below time it
        self.stream = torch.cuda.Stream()
        torch.cuda.set_stream(self.stream)
        assert self.stream.cuda_stream != 0

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        if self.dist is None:
            return [float(v) for v in vals]
        t = self.torch.tensor([float(v) for v in vals], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def engine(self, da, model, K, dtype, opt, sharded=True, seed=SEED):
        import barbay_b200 as bb
        rank, world = (self.rank, self.world) if sharded else (0, 1)
        eng = bb.Engine(da, model, n_samples=K, dtype=dtype, seed=seed, device=self.local_rank, rank=rank, world=world)
        if world > 1:
            self.wire(eng)
        eng.set_stream(self.stream.cuda_stream)
        eng.init_params(1)
        eng.set_optimizer(opt)
        return eng

    def wire(self, eng):
        """Connect the ranks' handles: CUDA IPC handles of the exchange buffers gathered over torch.distributed
        (bb_peer_attach: no communicator inside the library); BB_BENCH_NCCL=1 -> bb_comm_init (ncclCommInitRank)."""
        import barbay_b200 as bb
        if os.environ.get("BB_BENCH_NCCL"):
            uid = [bb.comm_unique_id() if self.rank == 0 else None]
            self.dist.broadcast_object_list(uid, src=0)
            eng.comm_init(uid[0])
        else:
            hs = [None] * self.world
            self.dist.all_gather_object(hs, eng.peer_handle())
            eng.peer_attach(hs)

    def time_steps(self, eng, steps, warmup):
        """(ms over `steps` steps, max over ranks; launches) -- CUDA events on the launching stream."""
        torch = self.torch
        eng.step(max(warmup, 3))
        self.barrier()
        l0 = eng.launch_count
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(self.stream)
        eng.step(steps)
        ev1.record(self.stream)
        self.barrier()
        ms = ev0.elapsed_time(ev1)
        (ms,) = self.max_over_ranks(ms)
        return ms, eng.launch_count - l0


def parity_check(cx: Ctx, K=2, n_barcodes=30_000, n_steps=4):
    """Sharded (all ranks) vs unsharded (rank 0) run of the same problem, same seed: fp64, `n_steps` ADVI steps.
    Returns {"mean_rel", "sd_rel", "elbo_rel", "ok"} on rank 0."""
    torch = cx.torch
    model, da = make_workload(CFG, 1, scale=n_barcodes / 1e6)
    out = {}
    eng = cx.engine(da, model, K, "f64", "decayed", sharded=True, seed=7)
    eng.step(n_steps)
    elbo = eng.step(1, elbo_trace=True)[0]
    m, s = eng.get_posterior()
    eng.close()
    t = torch.from_numpy(np.stack([m, s])).cuda()
    cx.dist.all_reduce(t)                       # every latent is owned (reported) by exactly one rank
    m, s = t.cpu().numpy()
    if cx.rank == 0:
        ref = cx.engine(da, model, K, "f64", "decayed", sharded=False, seed=7)
        ref.step(n_steps)
        elbo1 = ref.step(1, elbo_trace=True)[0]
        m1, s1 = ref.get_posterior()
        ref.close()
        out = {"mean_rel": float(np.max(np.abs(m - m1)) / np.max(np.abs(m1))),
               "sd_rel": float(np.max(np.abs(s - s1) / s1)),
               "elbo_rel": float(abs(elbo - elbo1) / abs(elbo1)),
               "what": f"{cx.world}-GPU sharded vs 1-GPU run, fitness_normal {n_barcodes} barcodes x 5, K={K}, fp64, "
                       f"{n_steps + 1} steps, same Philox seed"}
        out["ok"] = bool(out["mean_rel"] < 1e-9 and out["sd_rel"] < 1e-9 and out["elbo_rel"] < 1e-10)
    cx.barrier()
    return out


def parity_check_f32(cx: Ctx, K=K_MC, n_barcodes=120_000, n_steps=24):
    """The MEASURED path -- fp32 persistent step kernel with the in-kernel NVLink exchange -- sharded over all ranks
    against the same kernel on one GPU: one call of `n_steps` steps, same seed.  The column arithmetic is identical;
    only the order in which the per-thread fp32 partial sums enter the double totals differs with the tile split."""
    torch = cx.torch
    model, da = make_workload(CFG, 1, scale=n_barcodes / 1e6)
    out = {}
    eng = cx.engine(da, model, K, "f32", "decayed", sharded=True, seed=7)
    eng.step(n_steps)
    plane = eng.data_plane()
    m, s = eng.get_posterior()
    eng.close()
    t = torch.from_numpy(np.stack([m, s])).cuda()
    cx.dist.all_reduce(t)
    m, s = t.cpu().numpy()
    if cx.rank == 0:
        ref = cx.engine(da, model, K, "f32", "decayed", sharded=False, seed=7)
        ref.step(n_steps)
        m1, s1 = ref.get_posterior()
        ref.close()
        out = {"mean_rel": float(np.max(np.abs(m - m1)) / np.max(np.abs(m1))),
               "sd_rel": float(np.max(np.abs(s - s1) / s1)),
               "sd_rel_median": float(np.median(np.abs(s - s1) / s1)),
               "step_kernel": plane["step_kernel"], "persistent": plane["persistent"],
               "peer_exchange": plane["peer_exchange"],
               "what": f"{cx.world}-GPU sharded vs 1-GPU run of the measured fp32 path, fitness_normal {n_barcodes} "
                       f"barcodes x 5, K={K}, one call of {n_steps} steps, same Philox seed"}
        out["ok"] = bool(out["mean_rel"] < 1e-5 and out["sd_rel"] < 1e-4)
    cx.barrier()
    return out


def roofline_of(cx: Ctx, eng, steps_ms_per_step, n_prof, peak, peak_src, key, note=None):
    ms_tot, ms_p1, ms_p2 = eng.time_steps(n_prof)
    cx.barrier()
    alg = eng.algorithmic_bytes_per_step
    t_k = ms_p2 / n_prof * 1e-3
    t_p1 = ms_p1 / n_prof * 1e-3
    plane = eng.data_plane()
    fused = t_p1 < 0.1 * t_k
    kname = ("step_kernel<..., W=2> (packed fp32 gradient + optimiser update + next step's partial sums"
             + ("; persistent: in-kernel reduction / exchange / shared-latent phases included)" if plane["persistent"] else ")")
             ) if plane["step_kernel"] else ("pass2_kernel<..., FUSE=1>" if fused else "pass2_kernel (+ pass1_kernel)")
    traffic, tsrc = ncu_traffic(key)
    ach = alg / t_k / 1e9
    step_ach = alg / (steps_ms_per_step * 1e-3) / 1e9
    r = {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
         "traffic": traffic, "traffic_source": tsrc, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg,
         "kernel_us": t_k * 1e6, "pass1_us": t_p1 * 1e6, "step_us": steps_ms_per_step * 1e3,
         "kernel_share_of_step": t_k / (ms_tot / n_prof * 1e-3), "step_achieved": step_ach, "step_frac": step_ach / peak,
         "l2_resident": bool(alg < L2_BYTES)}
    if plane["persistent"]:
        # one launch covers up to persist_chunk steps: bytes and duration are those of the launch, the ratio is per step
        spl = max(1, min(int(plane["persist_chunk"]), int(n_prof)))
        r["steps_per_launch"] = spl
        r["algorithmic_bytes_per_step"] = alg
        r["algorithmic_bytes_per_launch"] = alg * spl
        r["launch_us"] = t_k * 1e6 * spl
        if traffic is not None:
            r["traffic_per_step"] = traffic
            r["traffic"] = traffic * spl
            r["traffic_note"] = ("ncu capture of a ONE-step launch of the same kernel (BB_PERSIST=0), scaled by "
                                 "steps_per_launch")
        r["kernel_us_note"] = ("persistent launch: kernel_us = launch duration / steps, i.e. it includes the in-kernel "
                               "reduction and shared-latent phases of every step")
    if note:
        r["note"] = note
    return r


def sub_measure(cx: Ctx, cfg, K, dtype, opt, peak, steps=200, sharded=False, warmup=10):
    """One extra configuration: step time, algorithmic GB/s of the whole step, fraction of peak.  sharded=False: on one
    GPU; sharded=True (multi-GPU runs): the same total problem sharded over all ranks (strong scaling), time = max over
    ranks, bytes and fraction per GPU."""
    model, da = make_workload(cfg)
    shape, n_units = count_shape(da)
    eng = cx.engine(da, model, K, dtype, opt, sharded=sharded)
    ms, _ = cx.time_steps(eng, steps, warmup)
    alg = eng.algorithmic_bytes_per_step
    plane = eng.data_plane()
    eng.close()
    step_s = ms / steps * 1e-3
    return {"model": model, "shape": list(shape), "mc_samples": K, "dtype": dtype, "optimizer": opt,
            "step_us": step_s * 1e6, "value": n_units * K / step_s, "unit": UNIT, "algorithmic_bytes": alg,
            "step_achieved": alg / step_s / 1e9, "step_frac": alg / step_s / 1e9 / peak,
            "step_kernel": plane["step_kernel"], "persistent": plane["persistent"]}


def run_ours(args, emit):
    import barbay_b200 as bb
    cx = Ctx()
    torch, world, rank = cx.torch, cx.world, cx.rank
    K = args.mc_samples
    peak, peak_src = measured_peak_gbs()

    parity = parity_check(cx) if world > 1 else None
    parity_f32 = parity_check_f32(cx) if world > 1 and args.dtype == "f32" else None

    sampler = ClockSampler(cx.local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    mult = world if args.scaling == "weak" else 1
    model, da = make_workload(CFG, mult)
    (T, B), n_cells = count_shape(da)
    units_step = n_cells * K
    eng = cx.engine(da, model, K, args.dtype, args.opt)
    plane = eng.data_plane()

    if sampler:
        sampler.wait_first()
    cx.barrier()
    t_wall0 = time.time()
    ms, launches = cx.time_steps(eng, args.steps, args.warmup)
    t_wall1 = time.time()
    value = units_step * args.steps / (ms * 1e-3)
    breakdown = eng.persist_stats() if plane["persistent"] else None

    n_prof = min(200, max(10, args.steps))
    key = f"cfg{CFG}_{args.dtype}_{args.opt}_k{K}_n{world}"
    roofline = roofline_of(
        cx, eng, ms / args.steps, n_prof, peak, peak_src, key,
        note="K=8 is issue-bound, not HBM-bound: both passes regenerate the Philox/Box-Muller noise (see DESIGN.md "
             "section 5 for the instruction budget); roofline_k1 is the same measurement at the reference's default "
             "samples_per_step=1")
    clocks = sampler.summary(t_wall0, time.time()) if sampler else None

    # ---- more lines on one GPU: reference default K=1, fp64 (the reference's arithmetic), the reference's default
    # optimiser, and the other BASELINE configurations
    extras = {}
    if world == 1 and not args.no_extras:
        eng1 = cx.engine(da, model, 1, args.dtype, args.opt, sharded=False)
        ms1, _ = cx.time_steps(eng1, n_prof, 10)
        r1 = roofline_of(cx, eng1, ms1 / n_prof, n_prof, peak, peak_src, f"cfg{CFG}_{args.dtype}_{args.opt}_k1_n1")
        r1["value"] = n_cells / (ms1 / n_prof * 1e-3); r1["mc_samples"] = 1
        extras["roofline_k1"] = r1
        eng1.close()
        if args.dtype != "f64":
            e64 = cx.engine(da, model, K, "f64", args.opt, sharded=False)
            ms64, _ = cx.time_steps(e64, n_prof, 10)
            r64 = roofline_of(cx, e64, ms64 / n_prof, n_prof, peak, peak_src, f"cfg{CFG}_f64_{args.opt}_k{K}_n1")
            extras["value_f64"] = units_step / (ms64 / n_prof * 1e-3)
            extras["roofline_f64"] = r64
            e64.close()
        # the pure ELBO-gradient evaluation (K draws + log-joint + analytic gradient, NO optimiser update;
        # bb_elbo_grad with the gradient left on the device: pass 1 + tail + pass 2 with the ELBO terms, gradients
        # written to their own arrays) -- CUDA events around `n_eg` calls, each ending with the 8-byte ELBO read-back
        eg = cx.engine(da, model, K, args.dtype, args.opt, sharded=False)
        for s_ in range(3):
            eg.elbo_grad(step=s_, want_grad=False)
        n_eg = 100
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(cx.stream)
        for s_ in range(n_eg):
            eg.elbo_grad(step=3 + s_, want_grad=False)
        ev1.record(cx.stream); torch.cuda.synchronize()
        eg_us = ev0.elapsed_time(ev1) / n_eg * 1e3
        eg_bytes = 4 * (4 if args.dtype != "f64" else 8) * eg.D + 4 * n_cells      # 4 w D + 4 T B R
        extras["elbo_grad_only"] = {"call_us": eg_us, "value": units_step / (eg_us * 1e-6), "unit": UNIT,
                                    "algorithmic_bytes": eg_bytes, "achieved": eg_bytes / (eg_us * 1e-6) / 1e9,
                                    "frac": eg_bytes / (eg_us * 1e-6) / 1e9 / peak,
                                    "what": "bb_elbo_grad(eps = NULL, grad = NULL): no update, gradient left on the device, ELBO read back"}
        eg.close()
        # throughput is data-independent (no data-dependent branch in any kernel): the same step on counts simulated
        # from another seed
        model_a, da_a = make_workload(CFG, seed_offset=1000)
        ea = cx.engine(da_a, model_a, K, args.dtype, args.opt, sharded=False)
        ms_a, _ = cx.time_steps(ea, 400, 20)
        extras["alt_seed"] = {"seed_offset": 1000, "step_us": ms_a / 400 * 1e3, "value": units_step / (ms_a / 400 * 1e-3), "unit": UNIT}
        ea.close(); del da_a
        if args.opt != "truncated":
            # fp32 TruncatedADAGrad rebuilds its window sums from the ring at the first four window wraps (n = 100 steps)
            # and at every eighth after them: time 800 steps past the fourth wrap, i.e. the steady state with one
            # rebuild in the timed region
            extras["truncated"] = sub_measure(cx, CFG, K, args.dtype, "truncated", peak, steps=800, warmup=410)
        extras["configs"] = {f"cfg{c}": sub_measure(cx, c, K, args.dtype, args.opt, peak) for c in (3, 4, 5)}

    # ---- weak scaling beside the strong headline (10^6 barcodes per GPU)
    scaling_weak = None
    if world > 1 and args.scaling == "strong" and not args.no_extras:
        wmodel, wda = make_workload(CFG, world)
        _, wcells = count_shape(wda)
        weng = cx.engine(wda, wmodel, K, args.dtype, args.opt)
        wsteps = min(args.steps, 500)
        wms, _ = cx.time_steps(weng, wsteps, 10)
        scaling_weak = {"value": wcells * K * wsteps / (wms * 1e-3), "unit": UNIT, "ms_per_step": wms / wsteps,
                        "barcodes": int(wcells // T), "steps": wsteps,
                        "breakdown_us": weng.persist_stats() if weng.data_plane()["persistent"] else None}
        weng.close()
        del wda
        # the other BASELINE configurations (round-1 kernels, exchange inside tail_kernel), sharded over the same ranks
        extras["configs_sharded"] = {f"cfg{c}": sub_measure(cx, c, K, args.dtype, args.opt, peak, sharded=True)
                                     for c in (3, 4, 5)}

    # ---- end to end through the public API with HOST buffers: one complete advi()-equivalent call
    # (bb_create: pack + H2D of counts / maps -> [comm] -> init -> optimiser -> n steps -> ELBO read-back ->
    # bb_get_posterior: D2H), timed on the host clock around the whole call (all ranks, max over ranks)
    n_e2e = args.e2e_steps
    cx.barrier()
    t0 = time.perf_counter()
    eng2 = bb.Engine(da, model, n_samples=K, dtype=args.dtype, seed=SEED, device=cx.local_rank, rank=rank, world=world)
    t_comm = 0.0
    if world > 1:
        tc0 = time.perf_counter()
        cx.wire(eng2)
        t_comm = time.perf_counter() - tc0       # gather of the ranks' IPC handles + peer mappings (no NCCL communicator)
    eng2.init_params(1)
    eng2.set_optimizer(args.opt)
    eng2.step(n_e2e)
    elbo_last = eng2.step(1, elbo_trace=True)            # the step's result read back (8 bytes)
    m, s = eng2.get_posterior()
    dt_e2e = time.perf_counter() - t0
    dt_e2e, t_comm = cx.max_over_ranks(dt_e2e, t_comm)
    n_tot = n_e2e + 1
    h2d = (n_cells * 4 + (n_cells + 2 * (B - da.n_neutral)) * 4) / world / n_tot     # int32 counts + layout maps per rank
    d2h = (2 * eng2.D * 8) / n_tot + 8.0 / n_tot                                    # posterior (m, sigma) + ELBO
    e2e = {"value": units_step * n_tot / dt_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "steps": n_tot, "seconds": dt_e2e,
           "what": "one complete advi()-equivalent call through the C ABI from HOST arrays on every rank: bb_create "
                   "(pack + H2D of counts / maps) + [comm init] + bb_init_params + bb_set_optimizer + bb_step + "
                   "ELBO read-back + bb_get_posterior (D2H); bytes are amortised over the steps of the call",
           "elbo_last": float(elbo_last[-1]), "posterior_finite": bool(np.isfinite(m).all())}
    if world > 1:
        e2e["comm_init_seconds"] = t_comm
        e2e["value_excluding_comm_init"] = units_step * n_tot / max(dt_e2e - t_comm, 1e-9)
    eng2.close()
    if rank == 0 and world == 1 and not args.no_extras:
        # streaming variant per the base contract: every step re-uploads that step's counts from pinned
        # host memory and reads the step's ELBO back
        cnt_host = torch.from_numpy(np.ascontiguousarray(np.asarray(da.bc_count).astype(np.int32))).pin_memory()
        cnt_dev = torch.empty_like(cnt_host, device="cuda")
        n_s = min(200, args.steps)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_s):
            cnt_dev.copy_(cnt_host, non_blocking=True)
            eng.step(1, elbo_trace=True)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        e2e["streaming"] = {"value": units_step * n_s / (t1 - t0), "unit": UNIT,
                            "h2d_bytes_per_step": int(cnt_host.numel() * 4), "d2h_bytes_per_step": 8 * (K + 1),
                            "steps": n_s, "what": "per step: pinned-host -> device copy of the count matrix + bb_step "
                            "(with ELBO terms) + ELBO read-back"}

    # ---- second end-to-end figure, from the TIDY DataFrame (what BarBay.vi.advi is handed): data_to_arrays (packing,
    # src/utils.jl:996-1033) + the fit + advi_to_df (src/utils.jl:1409-1462), each timed
    if rank == 0 and world == 1 and not args.no_extras:
        try:
            tidy = bb.synth.to_tidy_fast(da)
            t0 = time.perf_counter()
            da_p = bb.utils.data_to_arrays(tidy)
            t_pack = time.perf_counter() - t0
            t0 = time.perf_counter()
            eng3 = bb.Engine(da_p, model, n_samples=K, dtype=args.dtype, seed=SEED, device=cx.local_rank)
            eng3.init_params(1)
            eng3.set_optimizer(args.opt)
            n_df = min(n_e2e, 2000)
            eng3.step(n_df)
            m3, s3 = eng3.get_posterior()
            t_fit = time.perf_counter() - t0
            t0 = time.perf_counter()
            q = bb.utils.MeanFieldPosterior.build(m3, s3, eng3.layout.ranges_out)
            out_df = bb.utils.advi_to_df(tidy, q, eng3.layout.var_names, output=da_p)
            t_unpack = time.perf_counter() - t0
            eng3.close()
            e2e["from_dataframe"] = {
                "rows_in": int(len(tidy)), "rows_out": int(len(out_df)), "steps": n_df,
                "data_to_arrays_s": t_pack, "fit_s": t_fit, "advi_to_df_s": t_unpack,
                "value": units_step * n_df / (t_pack + t_fit + t_unpack), "unit": UNIT,
                "what": "tidy DataFrame -> data_to_arrays -> bb_create ... bb_step x steps ... bb_get_posterior -> advi_to_df "
                        "(posterior DataFrame), host wall clock"}
            del tidy, out_df
        except Exception as exc:                                   # reported, never fatal for the headline
            e2e["from_dataframe"] = {"error": repr(exc)[:200]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, cores, sample = cpu_port_rate(da, 15.0, K)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    if sampler:
        sampler.stop()
    if rank == 0:
        shard_bytes = eng.algorithmic_bytes_per_step
        if plane["persistent"] and world > 1:
            dp = "peer_ipc: in-kernel exchange of the step's sums over NVLink peer memory (persistent step kernel)"
        elif plane["peer_exchange"]:
            dp = "peer_ipc: exchange of the step's sums over NVLink peer memory inside the tail kernel"
        elif world > 1:
            dp = "nccl: one ncclAllReduce of the step's sums per step"
        else:
            dp = "none (single GPU)"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"BASELINE configs[1]: fitness_normal, {B} barcodes x {T} time points, {K} MC samples"
                       + (f" ({B // world} barcodes per GPU)" if world > 1 else ""),
                       "optimizer": "DecayedADAGrad" if args.opt == "decayed" else "TruncatedADAGrad(n=100)",
                       "l2": ("inputs_exceed_l2 (per-step working set > 126 MB, no flush needed)" if shard_bytes > L2_BYTES
                              else f"shard_fits_l2 ({shard_bytes / 1e6:.0f} MB per GPU per step < 126 MB L2: the strong-scaling "
                                   "shard is L2-resident by construction; HBM fraction is reported against algorithmic bytes)"),
                       "parallelism": f"barcode-sharded x{world}; {5 * T * K} doubles combined per step" if world > 1 else "single GPU",
                       "data_plane": dp,
                       "steps_per_launch": plane["persist_chunk"] if plane["persistent"] else 1},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        if breakdown:
            line["breakdown_us"] = breakdown
        if parity is not None:
            line["parity"] = parity
        if parity_f32 is not None:
            line["parity_f32"] = parity_f32
        if scaling_weak is not None:
            line["scaling_weak"] = scaling_weak
        line.update(extras)
        emit(line)
    eng.close()
    if cx.dist is not None:
        cx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--opt", default="decayed", choices=["decayed", "truncated"])
    ap.add_argument("--mc-samples", type=int, default=K_MC)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the K=1 / fp64 / TruncatedADAGrad / cfg3-5 / weak sub-lines")
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
    return run_ours(args, emit)


if __name__ == "__main__":
    main()
