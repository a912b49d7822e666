#!/usr/bin/env python
"""bench.py -- train rows/sec of the N x 128 least-squares fit (BASELINE.json `metric`).

One "step" = one complete fit of the resident rows through the C-ABI:
    b2_gram_reset -> b2_gram_accumulate (tcgen05 Gram kernel over every row of this rank's shard)
    -> b2_gram_allreduce (NCCL, N > 1 only) -> b2_solve (single-SM Cholesky, coefficients to the host)

Arms
    python bench.py [--gpus N --steps K --warmup W]          this repo (one process per GPU under torchrun)
    python bench.py --impl reference [...]                   the reference's own CPU path: scikit-learn
                                                             LinearRegression.fit (stage_1_train_model.py:105-106)
                                                             on a bounded sample, all host threads, rank 0 only

Keys of the JSON line follow the driver's contract; see DESIGN.md "Measurement".
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

D = 128
ROWS_N1 = 10_000_000       # BASELINE.json configs[1]
ROWS_PER_GPU_MULTI = 12_500_000  # BASELINE.json configs[2]: 100 M rows over 8 GPUs
METRIC = "train rows/sec (N x 128 least-squares fit)"
UNIT = "rows/s"


# ---------------------------------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_tensor_peak():
    """dense bf16 TFLOP/s (burst) of this pool's B200s, driver-written; fallback = the profiling recipe's figure"""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh).get("bf16_tflops", 1688.7))
    return 1688.7


def ncu_traffic_per_launch(workload_key: str):
    """dram bytes per launch of the Gram kernel from the committed ncu capture, if one matches."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh).get(workload_key)
    return None


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region, polled in-process through NVML every ~2 ms
    (the timed region of a 20-step run is ~20 ms -- too short for an `nvidia-smi -lms` subprocess to start).
    Falls back to the B200_PROFILING.md nvidia-smi recipe when pynvml is unavailable."""
    _REASONS = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20),
                ("hw_thermal_slowdown", 0x40), ("hw_power_brake_slowdown", 0x80))

    def __init__(self, device: int):
        self.device, self.samples, self.stop_flag, self.thread, self.h = device, [], False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # torchrun may remap devices through CUDA_VISIBLE_DEVICES; NVML indexes physical devices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[device]) if vis and vis.split(",")[device].isdigit() else device
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.samples.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                                     nv.nvmlDeviceGetCurrentClocksEventReasons(self.h),
                                     nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0))
            except Exception:
                break
            time.sleep(0.002)

    def start(self):
        if self.nv is None:
            return
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def stop(self) -> dict:
        if self.nv is None:
            return self._smi_once()
        self.stop_flag = True
        self.thread.join(timeout=2)
        if not self.samples:
            return self._smi_once()
        sm = sorted(s[0] for s in self.samples)
        mask = 0
        for s in self.samples:
            mask |= int(s[1])
        reasons = [name for name, bit in self._REASONS if mask & bit]
        try:
            mx = self.nv.nvmlDeviceGetMaxClockInfo(self.h, self.nv.NVML_CLOCK_SM)
        except Exception:
            mx = None
        return {"sm_mhz": float(sm[len(sm) // 2]), "sm_max_mhz": float(mx) if mx else None, "samples": len(sm),
                "power_w_max": max(s[2] for s in self.samples), "reasons": reasons, "how": "NVML poll, 2 ms"}

    def _smi_once(self) -> dict:
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                  str(self.device)], capture_output=True, text=True, timeout=10).stdout.strip()
            f = [x.strip() for x in out.split(",")]
            names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
            return {"sm_mhz": float(f[0]), "sm_max_mhz": float(f[1]), "samples": 1, "power_w_max": float(f[2]),
                    "reasons": [n for n, v in zip(names, f[3:7]) if v.lower().startswith("active")],
                    "how": "nvidia-smi, one sample right after the timed region"}
        except Exception as exc:  # pragma: no cover
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"unavailable: {exc}"]}


# ---------------------------------------------------------------------------------------------------
def sklearn_fit_rows_per_s(rows: int, repeats: int = 1, seed: int = 1234):
    """The reference's fit call on `rows` x 128 fp32 rows drawn from the reference DGP; all host threads."""
    from sklearn.linear_model import LinearRegression
    from oracle import ols_oracle as orc
    X, y = orc.generate_dataset(rows, D, seed=seed, dtype=np.float32)
    best = float("inf")
    coef = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        reg = LinearRegression(fit_intercept=True).fit(X, y)
        best = min(best, time.perf_counter() - t0)
        coef = reg.coef_
    return rows / best, best, coef


def host_threads() -> int:
    try:
        from threadpoolctl import threadpool_info
        n = [p.get("num_threads", 0) for p in threadpool_info() if p.get("user_api") == "blas"]
        if n:
            return int(max(n))
    except Exception:
        pass
    return os.cpu_count() or 1


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import sklearn  # noqa: F401  (the reference's dependency; pinned 0.24.0 upstream, 1.9.0 in this image)
    try:   # torchrun exports OMP_NUM_THREADS=1: give the CPU arm every host thread BLAS will take
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass
    t_cal = sklearn_fit_rows_per_s(32_768)[1]
    budget = 150.0 / max(args.steps + args.warmup, 1)
    rows = int(min(1_000_000, max(32_768, 32_768 * budget / max(t_cal, 1e-3) * 0.5)))
    rows = (rows // 32_768) * 32_768
    from sklearn.linear_model import LinearRegression
    from oracle import ols_oracle as orc
    X, y = orc.generate_dataset(rows, D, seed=1234, dtype=np.float32)
    for _ in range(args.warmup):
        LinearRegression(fit_intercept=True).fit(X, y)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        LinearRegression(fit_intercept=True).fit(X, y)
    dt = time.perf_counter() - t0
    value = rows * args.steps / dt
    cores = host_threads()
    sample = f"{rows} x {D} fp32 rows of the same synthetic distribution per step (bounded sample of the workload)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, "f32"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "what": "sklearn.linear_model.LinearRegression(fit_intercept=True).fit "
                                 "(stage_1_train_model.py:105-106; LAPACK gelsd), called as the oracle port does"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_config(n_gpus: int, x_kind: str) -> dict:
    rows = ROWS_N1 if n_gpus == 1 else ROWS_PER_GPU_MULTI
    return {"workload": (f"{rows * n_gpus} rows x {D} features ({rows} per GPU, row-sharded), X {x_kind} + y fp32 "
                         f"resident in HBM; BASELINE.json configs[{1 if n_gpus == 1 else 2}]"),
            "rows_per_gpu": rows, "features": D, "x_storage": x_kind,
            "parallelism": f"row-shard x{n_gpus}, one all-reduce of the (D+2)^2 fp64 statistic per fit (peer-memory one-shot "
                           f"exchange over NVLink; NCCL all-reduce when B2_NO_P2P=1)",
            "l2": "inputs larger than L2 (5.2 GB per pass vs 126 MB)"}


# ---------------------------------------------------------------------------------------------------
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--x-dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--rows", type=int, default=0, help="rows per GPU (default: BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--precision", default="split", choices=["split", "bf16"],
                    help="tensor-core operand precision: bf16 hi+lo split (default) or a single bf16 operand")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference(args)
        return

    import bodywork_mlops_demo_b200 as b2

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N > 1 with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N")
        args.gpus = world

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist  # rendezvous / barrier / max-over-ranks only (gloo, CPU tensors)
        dist.init_process_group("gloo", init_method="env://", rank=rank, world_size=world)

    ctx = b2.Context(local_rank)
    if world > 1:
        import torch
        uid = [b2.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        # NCCL may print its version banner on stdout; stdout carries exactly one JSON line, so park fd 1 on stderr
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            ctx.comm_init(world, rank, uid[0])
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
        if os.environ.get("B2_NO_P2P") != "1" and world <= 8:
            # one-shot peer-memory exchange of S (NVLink stores + flags) instead of an NCCL launch per step
            try:
                mine = ctx.comm_p2p_export()
            except Exception as exc:   # e.g. CUDA IPC unavailable in this container
                print(f"[bench] rank {rank}: peer-memory exchange unavailable ({exc})", file=sys.stderr)
                mine = None
            handles = [None] * world
            dist.all_gather_object(handles, mine)          # every rank takes part in both collectives
            ok = all(h is not None for h in handles)
            if ok:
                try:
                    ctx.comm_p2p_attach(world, rank, handles)
                except Exception as exc:
                    print(f"[bench] rank {rank}: attaching peer buffers failed ({exc}); using NCCL", file=sys.stderr)
                    ok = False
            oks = [None] * world
            dist.all_gather_object(oks, ok)
            if not all(oks):
                ctx.comm_p2p_detach()

    rows = args.rows or (ROWS_N1 if world == 1 else ROWS_PER_GPU_MULTI)
    kind = args.x_dtype
    X, y = ctx.synth(rows, D, seed=1234, row_offset=rank * rows, kind=kind)
    ctx.set_kernel(b2.KERNEL_TCGEN05)
    ctx.set_precision(b2.PRECISION_BF16 if args.precision == "bf16" else b2.PRECISION_SPLIT)
    ctx.sync()

    def step():
        ctx.gram_reset(D)
        ctx.gram_accumulate(X, y)
        ctx.gram_allreduce()
        return ctx.solve()

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()

    for _ in range(args.warmup):
        coef, b0 = step()
    ctx.last_kernel_ms()
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    ctx.timer_start()
    t_host0 = time.perf_counter()
    for _ in range(args.steps):
        coef, b0 = step()
    ms = ctx.timer_stop()
    t_host = time.perf_counter() - t_host0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    kernel_ms, kernel_launches = ctx.last_kernel_ms()
    launches = ctx.launch_count() - launches0
    if dist is not None:
        import torch
        t = torch.tensor([ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())

    total_rows = rows * world
    value = total_rows * args.steps / (ms * 1e-3)
    bytes_per_row = D * (4 if kind == "f32" else 2) + 4
    peak, peak_src = measured_peaks()
    gram_ms = kernel_ms / max(kernel_launches, 1)
    achieved = rows * bytes_per_row / (gram_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "gram_tc_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src,
                "mma_tflops_issued": rows * 2.0 * 128 * 144 * (2 if args.precision == "split" else 1) / gram_ms / 1e9
                                     if D == 128 else None,
                "algorithmic_bytes_per_row": bytes_per_row, "rows_per_launch": rows,
                "kernel_ms_avg": gram_ms, "kernel_share_of_step": gram_ms / (ms / args.steps),
                "traffic": ncu_traffic_per_launch(f"{kind}_{rows}x{D}")}

    # ---- e2e: the public estimator API on HOST (pinned) rows; H2D inside the timed region ----------------
    e2e = None
    if not args.no_e2e:
        e2e_rows = rows
        Xp = ctx.pinned((e2e_rows, D), np.float32 if kind == "f32" else np.uint16)
        yp = ctx.pinned((e2e_rows,), np.float32)
        b2.native._check(b2.native.load().b2_copy_d2h(ctx._h, Xp.ptr, X.ptr, X.nbytes), "d2h X")
        b2.native._check(b2.native.load().b2_copy_d2h(ctx._h, yp.ptr, y.ptr, y.nbytes), "d2h y")
        est = b2.B200LinearRegression(ctx=ctx)
        ctx.set_kernel(b2.KERNEL_AUTO)
        e2e_steps = max(2, min(args.steps, 5))
        est.fit(Xp.array, yp.array, with_spectrum=False)   # warm-up (allocates the staging ring)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            est.fit(Xp.array, yp.array, with_spectrum=False)
        dt = time.perf_counter() - t0
        barrier()
        if dist is not None:
            import torch
            t = torch.tensor([dt], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": total_rows * e2e_steps / dt, "unit": UNIT, "steps": e2e_steps,
               "h2d_bytes_per_step": int(e2e_rows * bytes_per_row), "d2h_bytes_per_step": int((D + 1) * 8 + 8 * 4),
               "api": "B200LinearRegression.fit(X_host_pinned, y_host_pinned) -> b2_gram_accumulate(B2_MEM_HOST)",
               "coef_linf_vs_resident": float(np.max(np.abs(est.coef_ - coef)))}
        Xp.free(); yp.free()

    # ---- other operand / storage variants of configs[1] (N = 1 only; kernel + whole-fit rates, same timing rules) --
    variants = None
    if world == 1 and not args.no_variants:
        variants = {}
        for vk, vprec in ((kind, "bf16"), ("bf16" if kind == "f32" else "f32", "split"),
                          ("bf16" if kind == "f32" else "f32", "bf16")):
            Xv, yv = (X, y) if vk == kind else ctx.synth(rows, D, seed=1234, kind=vk)
            ctx.set_kernel(b2.KERNEL_TCGEN05)
            ctx.set_precision(b2.PRECISION_BF16 if vprec == "bf16" else b2.PRECISION_SPLIT)
            for _ in range(3):
                ctx.gram_reset(D); ctx.gram_accumulate(Xv, yv); ctx.solve()
            ctx.last_kernel_ms()
            ctx.sync(); ctx.timer_start()
            for _ in range(10):
                ctx.gram_reset(D); ctx.gram_accumulate(Xv, yv); cv, _b = ctx.solve()
            vms = ctx.timer_stop() / 10
            kms, kl = ctx.last_kernel_ms()
            bpr = D * (4 if vk == "f32" else 2) + 4
            variants[f"x_{vk}_operands_{'bf16x1' if vprec == 'bf16' else 'bf16x2'}"] = {
                "fit_rows_per_s": rows / vms * 1e3, "gram_kernel_ms": kms / max(kl, 1),
                "frac_of_hbm_peak": rows * bpr / (kms / max(kl, 1)) / 1e6 / peak,
                # flops as issued to the MMA: 2 * 128 * 144 per row and operand (hi, and lo in split mode)
                "mma_tflops_issued": rows * 2.0 * 128 * 144 * (1 if vprec == "bf16" else 2) / (kms / max(kl, 1)) / 1e9,
                "frac_of_bf16_tensor_peak": rows * 2.0 * 128 * 144 * (1 if vprec == "bf16" else 2) / (kms / max(kl, 1))
                                            / 1e9 / measured_tensor_peak(),
                "coef_linf_vs_headline_fit": float(np.max(np.abs(cv - coef)))}
            ctx.set_precision(b2.PRECISION_BF16 if args.precision == "bf16" else b2.PRECISION_SPLIT)
            if vk != kind:
                Xv.free(); yv.free()

    # ---- companion kernel: hold-out scoring + metrics over the same resident rows (stage_1...:107, 79-90) ----------
    companion = None
    if world == 1 and not args.no_variants:
        ctx.set_kernel(b2.KERNEL_AUTO)
        for _ in range(3):
            ctx.score(X, coef, float(b0), y=y, want_yhat=False)
        ctx.sync(); ctx.timer_start()
        for _ in range(10):
            _yh, sstats = ctx.score(X, coef, float(b0), y=y, want_yhat=False)
        sms = ctx.timer_stop() / 10
        sbpr = D * (4 if kind == "f32" else 2) + 4
        companion = {"what": "b2_score: X.coef + intercept fused with the ten metric reductions, resident rows, no yhat write",
                     "ms": sms, "rows_per_s": rows / sms * 1e3, "bytes_per_row": sbpr,
                     "frac_of_hbm_peak": rows * sbpr / sms / 1e6 / peak,
                     "note": "timed with CUDA events around 10 calls; includes the 80-byte D2H of the statistics per call",
                     "r_squared": float(1.0 - sstats[1] / max(sstats[3] - sstats[2] ** 2 / max(sstats[5], 1.0), 1e-300))}

    # ---- CPU baseline: sklearn on the host cores, bounded sample, rank 0 at N = 1 only --------------------
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample_rows = 1_000_000
        v, secs, _ = sklearn_fit_rows_per_s(sample_rows)
        cpu = {"value": v, "unit": UNIT, "cores": host_threads(), "kind": "port", "seconds": secs,
               "sample": f"{sample_rows} x {D} fp32 rows, one LinearRegression(fit_intercept=True).fit "
                         f"(stage_1_train_model.py:105-106), all BLAS threads"}
        # coefficient parity on rows both sides see: first 200k rows of the device buffer
        from sklearn.linear_model import LinearRegression
        m = 200_000
        Xh = X.to_host()[:m]
        yh = y.to_host()[:m]
        Xf = Xh.astype(np.float64) if kind == "f32" else b2.native.from_bf16_bits(Xh).astype(np.float64)
        reg = LinearRegression().fit(Xf, yh.astype(np.float64))
        sub = b2.B200LinearRegression(ctx=ctx)
        ctx.set_kernel(b2.KERNEL_TCGEN05)
        Xs, ys = ctx.to_device(Xh, kind), ctx.to_device(yh)
        sub.fit(Xs, ys, with_spectrum=False)
        parity = {"rows": m, "coef_linf_vs_sklearn_fp64": float(np.max(np.abs(sub.coef_ - reg.coef_))),
                  "intercept_abs_err": float(abs(sub.intercept_ - reg.intercept_)), "tolerance": 1e-4}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": ("bf16x2 (hi+lo)" if args.precision == "split" else "bf16x1") + " MMA operands, f32 accumulate, f64 fold+solve",
            "data": "synthetic (device Philox, reference DGP: X~U(0,100), y=1+0.5*sum(X)+10*eps)",
            "config": workload_config(world, kind), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "variants": variants,
            "companion_score": companion,
            "host_wall_ms_per_step": 1e3 * t_host / args.steps,
            "coef_head": [float(c) for c in coef[:3]], "intercept": float(b0),
        }
        print(json.dumps(out))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
