#!/usr/bin/env python
"""Benchmark of the KLHR hot path (BASELINE.json metric: chain-draws/sec, ESS/sec, % roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    torchrun --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): stan/ill-normal, D = 100 diagonal Gaussian, 65 536 chains
per GPU, Gaussian line family, fp64, reference warm-up schedule [50, 150, 350, 1000] run
before the timed region.  One "step" = every chain advanced by --draws-per-step KLHR draws
(one `klhr_run` launch).  N > 1 shards chains over ranks (weak scaling, no collective in the
step; the only NCCL traffic is the pooled-adaptation allreduce at window closures, which
happen in the untimed adaptation phase).

The timed number `value` has chain state resident in HBM; `e2e` re-uploads the start state
from pinned host memory and reads the final state back every step through the public sampler
API.  `--impl reference` times the reference's own CPU implementation of the same step (the
SciPy-BFGS single-chain port in oracle/ref_port.py, one process per chain on all host cores,
like reference run_experiments:27) -- /root/reference itself is pure Python and is not
present on the GPU box, so `cpu_baseline.kind` is "port".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "chain_draws_per_sec"
UNIT = "chain-draws/s"
SEED = 20261018


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--chains", type=int, default=65_536, help="chains per GPU")
    ap.add_argument("--dim", type=int, default=100)
    ap.add_argument("--draws-per-step", type=int, default=1000)
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--adapt-warmup", type=int, default=1000, help="sampler warm-up draws (untimed)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ess", action="store_true")
    ap.add_argument("--ref-draws-per-step", type=int, default=400)
    ap.add_argument("--ref-procs", type=int, default=0, help="0 = all host cores")
    return ap.parse_args()


# =============================================================================== CPU reference arm
def _ref_worker(conn, dim, seed, adapt_warmup):
    """One OS process = one chain of the reference algorithm (oracle port, SciPy BFGS)."""
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"                     # one core per chain; avoids BLAS oversubscription
    import numpy as np  # noqa: F401
    from oracle.bsmodel import BSModel
    from oracle.ref_port import ChainSampler
    model = BSModel(stan_file="stan/ill-normal.stan", data={"D": dim})
    s = ChainSampler(model, family="gauss", seed=seed, warmup=adapt_warmup)
    for _ in range(adapt_warmup):
        s.draw()
    conn.send("ready")
    while True:
        n = conn.recv()
        if n <= 0:
            break
        t0 = time.perf_counter()
        for _ in range(n):
            s.draw()
        conn.send(time.perf_counter() - t0)
    conn.close()


class RefPool:
    def __init__(self, procs, dim, adapt_warmup):
        import multiprocessing as mp
        ctx = mp.get_context("spawn")
        self.conns, self.procs = [], []
        for i in range(procs):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_ref_worker, args=(b, dim, SEED + i, adapt_warmup), daemon=True)
            p.start()
            self.conns.append(a)
            self.procs.append(p)
        for c in self.conns:
            assert c.recv() == "ready"

    def step(self, n):
        t0 = time.perf_counter()
        for c in self.conns:
            c.send(n)
        for c in self.conns:
            c.recv()
        return time.perf_counter() - t0

    def close(self):
        for c in self.conns:
            c.send(0)
        for p in self.procs:
            p.join(timeout=10)


def cpu_baseline(dim, draws, procs, adapt_warmup, steps=1, warmup=0):
    procs = procs or (os.cpu_count() or 1)
    pool = RefPool(procs, dim, adapt_warmup)
    try:
        for _ in range(warmup):
            pool.step(max(1, draws // 4))
        times = [pool.step(draws) for _ in range(steps)]
    finally:
        pool.close()
    total = sum(times)
    return dict(value=procs * draws * steps / total, unit=UNIT, cores=procs, kind="port",
                sample=f"{procs} processes x 1 chain (oracle/ref_port.py, SciPy BFGS, analytic NumPy model "
                       f"shim; BridgeStan absent), {adapt_warmup} warm-up draws untimed, then "
                       f"{steps} x {draws} draws each, wall clock"), total / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, sec_per_step = cpu_baseline(args.dim, args.ref_draws_per_step, args.ref_procs, args.adapt_warmup,
                                    steps=args.steps, warmup=args.warmup)
    line = {
        "metric": METRIC, "value": cb["value"], "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"stan/ill-normal D={args.dim} KLHR Gaussian line fit, one chain per host core",
                   "chains": cb["cores"], "draws_per_step": args.ref_draws_per_step},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# =============================================================================== clocks
class ClockSampler:
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(gpu_index)], stdout=self.tmp, stderr=subprocess.DEVNULL)
            import atexit
            atexit.register(self._kill)                  # never leave the sampler behind if the bench dies
        except OSError:
            pass

    def _kill(self):
        if self.proc is not None and self.proc.poll() is None:
            self.proc.kill()

    def stop(self, t_begin=None, t_end=None):
        """Median SM clock and active throttle reasons over samples inside [t_begin, t_end]
        (time.time() stamps of the timed region); all samples if the window holds none."""
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        rows = [r.split(",") for r in Path(self.tmp.name).read_text().splitlines() if r.count(",") >= 7]
        os.unlink(self.tmp.name)
        sm, reasons = [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        mx = None
        def stamp(s):
            try:
                return datetime.datetime.strptime(s.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
            except ValueError:
                return None
        if t_begin is not None:
            inside = [r for r in rows if (stamp(r[0]) or 0) >= t_begin - 0.05 and (stamp(r[0]) or 0) <= t_end + 0.05]
            out["samples_total"] = len(rows)
            rows = inside or rows
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if r[4 + k].strip().lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


# =============================================================================== B200 arm
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import klhr_b200 as kb
    from klhr_b200.diagnostics import chain_summary

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    # started first: nvidia-smi can take a second to spin up on a fresh box, and only samples stamped inside the
    # timed region are used
    clocks = ClockSampler(local) if rank == 0 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    B, D, S, K, W = args.chains, args.dim, args.draws_per_step, args.steps, args.warmup
    rb = 8 if dtype == torch.float64 else 4

    model = kb.BSModel(stan_file="stan/ill-normal.stan", data={"D": D}, device=dev)
    sampler = kb.KLHR(model, seed=SEED, chains=B, warmup=args.adapt_warmup, windowsize=50, windowscale=2,
                      dtype=dtype, device=dev)
    sampler.run(args.adapt_warmup)            # adaptation phase: windows close at 50,150,350,1000 (+ allreduce)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    # ---------------------------------------------------------------- resident arm
    for _ in range(W):
        sampler.run(S)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    wall0 = time.time()
    for k in range(K):
        flush.zero_()                          # L2 flush between timed iterations (state is 52 MB < L2)
        evs[k][0].record()
        sampler.run(S)                         # ONE klhr_run launch: B chains x S draws
        evs[k][1].record()
    torch.cuda.synchronize()
    barrier()
    wall = time.perf_counter() - t0
    step_ms = [a.elapsed_time(b) for a, b in evs]
    dev_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dev_ms, op=dist.ReduceOp.MAX)
    total_ms = float(dev_ms.item())
    clk = clocks.stop(wall0, time.time()) if clocks else None
    value = world * B * S * K / (total_ms * 1e-3)

    # ---------------------------------------------------------------- end-to-end arm (host buffers)
    # Every step: H2D of the step's start state from pinned host memory, S draws through the public
    # sampler API, D2H of the final state and the accept counts.  Two device state buffers and three
    # streams pipeline the copies of step k +- 1 under the kernel of step k (steps are independent
    # jobs: each starts from the host start state), so PCIe time overlaps compute.
    start_host = torch.empty(B, D, dtype=dtype).pin_memory()
    start_host.copy_(sampler._theta)
    out_host = [torch.empty(B, D, dtype=dtype).pin_memory() for _ in range(2)]
    acc_host = [torch.empty(B, dtype=torch.int64).pin_memory() for _ in range(2)]
    bufs = [sampler._theta, torch.empty_like(sampler._theta)]
    s_in, s_run, s_out = torch.cuda.Stream(dev), torch.cuda.current_stream(dev), torch.cuda.Stream(dev)

    def e2e_steps(n):
        ev_in = [torch.cuda.Event() for _ in range(n)]
        ev_run = [torch.cuda.Event() for _ in range(n)]
        ev_out = [torch.cuda.Event() for _ in range(n)]
        for k in range(n):
            buf = bufs[k % 2]
            with torch.cuda.stream(s_in):
                if k >= 2:
                    s_in.wait_event(ev_out[k - 2])               # buffer free again once step k-2 was read back
                buf.copy_(start_host, non_blocking=True)         # H2D of the step's start state
                ev_in[k].record(s_in)
            s_run.wait_event(ev_in[k])
            sampler.swap_state(buf)
            sampler.run(S)                                       # public sampler API, one launch
            ev_run[k].record(s_run)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_run[k])
                out_host[k % 2].copy_(buf, non_blocking=True)    # D2H of the step's result
                acc_host[k % 2].copy_(sampler._accept_count, non_blocking=True)
                ev_out[k].record(s_out)
        torch.cuda.synchronize()

    e2e_steps(min(W, 2))
    barrier()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_steps(K)
    e1.record()
    torch.cuda.synchronize()
    barrier()
    sampler.swap_state(bufs[0])
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * B * S * K / (float(e2e_ms.item()) * 1e-3)

    # ---------------------------------------------------------------- ESS per draw (diagnostic pass)
    ess = None
    if not args.no_ess:
        S_ess = max(S, 2000)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        s1, s2 = sampler.run(S_ess, chain_stats=True)
        a1.record()
        torch.cuda.synchronize()
        summ = chain_summary(s1, s2, S_ess)
        truth = torch.arange(1, D + 1, dtype=torch.float64, device=dev) ** 2 / D      # var_i = i^2 / D
        zmean = (summ["mean"].abs() / summ["mcse_mean"]).max()
        zvar = ((summ["var"] - truth).abs() / summ["mcse_var"]).max()
        ess_min = float(summ["ess"].min())
        ms = a0.elapsed_time(a1)
        # ESS per draw is a property of the chains (both kernels draw the same streams); the pass that
        # measures it needs per-chain sums and therefore runs on the accumulating octet kernel, so
        # ESS/sec of the production path = ESS per chain-draw x the timed chain-draws/sec
        ess = {"min_ess_per_sec": (ess_min / (B * S_ess)) * value, "min_ess_per_draw": ess_min / (B * S_ess),
               "min_ess_per_sec_of_stats_pass": world * ess_min / (ms * 1e-3),
               "draws_per_chain": S_ess, "estimator": "between-chain variance of chain means, B independent chains",
               "max_abs_z_mean": float(zmean), "max_abs_z_var": float(zvar),
               "acceptance": sampler.acceptance_probability}

    if rank == 0:
        peaks_path = ROOT / "MEASURED_PEAKS.json"
        if peaks_path.exists():
            peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        bytes_per_draw = 2 * D * rb + 16                        # theta read + write + scalars
        launch_ms = total_ms / K                                 # one step-kernel launch per step
        achieved = bytes_per_draw * B * S / (launch_ms * 1e-3) / 1e9
        traffic = None
        tp = ROOT / "profiles" / "ncu_traffic.json"
        if tp.exists():
            traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
        info = kb.launch_info(model, sampler._fit, dtype=dtype, free_running=True, accumulate=False, device=dev)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"stan/ill-normal D={D} diagonal Gaussian, {B} chains/GPU, KLHR Gaussian "
                                   f"line fit, {S} draws per step (one launch)",
                       "chains_per_gpu": B, "global_chains": world * B, "dim": D, "draws_per_step": S,
                       "adapt_warmup_draws": args.adapt_warmup, "windows": sampler._windowedadaptation.closures,
                       "parallelism": f"chains sharded over {world} GPU(s), no collective in the step",
                       "l2": "256 MB memset between timed steps (state 52 MB < 126 MB L2); excluded from "
                             "ms_per_step by per-step CUDA events", "seed": SEED,
                       "launch": info, "wall_s_incl_flush": wall},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * D * rb,
                    "d2h_bytes_per_step": B * D * rb + B * 8,
                    "api": "KLHR.run(draws_per_step) from a pinned host start state, final state + accept counts "
                           "read back every step; copies of neighbouring steps overlap the kernel (2 buffers, 3 streams)"},
            "gpu_launches": K,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": ("klhr::tile_kernel (csrc/klhr_tile.cuh)" if info["threads"] == 32 else
                                    "klhr::step_kernel (csrc/klhr_step.cuh)"),
                         "algorithmic_bytes_per_chain_draw": bytes_per_draw,
                         "note": f"achieved = SURVEY 8(d) algorithmic bytes x chain-draws / launch time; {S} draws are "
                                 "fused per launch and theta (52 MB) stays L2-resident between draws, so DRAM "
                                 "traffic (ncu) is far below the algorithmic bytes"},
            "ess": ess,
        }
        if world == 1 and not args.no_cpu_baseline:
            cb, _ = cpu_baseline(D, 2000, args.ref_procs, args.adapt_warmup)
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
