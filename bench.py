#!/usr/bin/env python
"""bench.py -- the hot path on the configuration BASELINE.json's metric is quoted on:
fp64 SpAMM C = A*B (tau = 1e-6) of two N = 65536 exponential-decay matrices, leaf 64 (SURVEY 8d cfg-2 law,
lambda = 0.01), at 1/2/4/8 B200 (strong scaling: C sharded by block rows, B halo exchanged over NCCL).

One JSON line on stdout (rank 0).  `value` = leaf-GEMM FP64 TFLOP/s over the WHOLE multiply (2 b^3 P / time per
multiply, task-list build included, inputs resident in HBM); `ms_per_step` = time per multiply; `e2e` = the same
through the C ABI with HOST buffers (pinned H2D of A and B tiles, norm refresh, SpAMM, D2H of all C tiles).
`--impl reference` times the unmodified reference (oracle/_ref, OpenMP on the host cores) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
from hierarchical_block_sparse_lib_b200 import _capi  # noqa: E402

WORKLOAD = dict(n=65536, b=64, lam=0.01, tau=1e-6, eps=1e-12, seeds=(1, 2))
CPU_SAMPLE_N = int(os.environ.get("HBSM_CPU_SAMPLE_N", "8192"))   # leading principal block of the same matrices, same law
FP64_PEAK_FILE = os.path.join(ROOT, "profiles", "r01_peak_fp64.json")


def fp64_peak():
    try:
        d = json.loads(open(FP64_PEAK_FILE).read().strip().splitlines()[0])
        return float(d["dmma884_sustained_tflops"]), "measured (tools/peak_fp64.cu on this pool's B200, profiles/r01_peak_fp64.json; MEASURED_PEAKS.json has no FP64 entry)"
    except Exception:
        return 37.0, "fallback (nominal B200 FP64)"


class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region (nvidia-smi, every 200 ms)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference_cpu(steps, warmup, n_sample=CPU_SAMPLE_N):
    """The reference's own OpenMP implementation (oracle/_ref = unmodified header compiled in place) on the host
    cores, on the leading n_sample x n_sample block of the workload's matrices.  Returns (tflops, ms, info)."""
    cores = host_threads()
    if "HBSM_REF_THREADS" in os.environ:
        os.environ["OMP_NUM_THREADS"] = os.environ["HBSM_REF_THREADS"]
    elif "TORCHELASTIC_RUN_ID" in os.environ or "OMP_NUM_THREADS" not in os.environ:
        # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm is ONE process that gets all host threads
        os.environ["OMP_NUM_THREADS"] = str(cores)
    cores = int(os.environ["OMP_NUM_THREADS"])      # what the run really uses (reported as cpu_baseline.cores)
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    from oracle import pyoracle as po
    from hierarchical_block_sparse_lib_b200 import generators as G
    w = WORKLOAD
    W = G.decay_width(w["lam"], w["eps"])
    kind = "reference" if os.path.exists(po.REF_SO) else "port"
    cls = po.RefMatrix if kind == "reference" else po.OrcMatrix
    mats = []
    for seed in w["seeds"]:
        r, c, v = G.decay_coo(n_sample, w["lam"], min(W, n_sample - 1), seed)
        mats.append(po.from_coo(cls, w["b"], n_sample, n_sample, r, c, v))   # norms refreshed (updated=true)
        del r, c, v
    times, nm = [], 0
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        Cm, nm, nb, _ = cls.product(mats[0], 0, mats[1], 0, spamm=True, tau=w["tau"])
        dt = time.perf_counter() - t0
        del Cm
        if i >= warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    tflops = 2.0 * w["b"] ** 3 * nm / (ms * 1e-3) / 1e12
    blas = po.RefMatrix.blas_kind() if kind == "reference" else "builtin loops"
    info = {"kind": kind, "cores": cores if kind == "reference" else 1,
            "sample": "leading %dx%d block of the same decay matrices (lambda=%g, tau=%g, b=%d): %d leaf products per "
                      "multiply, whole spamm() call (reserve+symbolic+numeric), OpenMP over hash buckets, BLAS=%s"
                      % (n_sample, n_sample, w["lam"], w["tau"], w["b"], nm, os.path.basename(blas)),
            "ms_per_multiply": ms, "products": nm}
    return tflops, ms, info


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    tflops, ms, info = run_reference_cpu(args.steps, args.warmup)
    line = {"metric": "spamm_fp64_leaf_tflops", "value": tflops, "unit": "TFLOP/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": tflops, "unit": "TFLOP/s", "cores": info["cores"], "kind": info["kind"],
                             "sample": info["sample"]},
            "e2e": {"value": tflops, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    w = WORKLOAD
    return {"workload": "fp64 SpAMM C=A*B, exponential-decay a_ij=(0.5+0.5u)exp(-%g|i-j|) truncated at 1e-12, N=%d, "
                        "leaf %d, tau=%g (BASELINE configs[1] law at the metric's N=65536)" % (w["lam"], w["n"], w["b"], w["tau"]),
            "n": w["n"], "leaf": w["b"], "lambda": w["lam"], "tau": w["tau"],
            "sharding": "single GPU" if n_gpus == 1 else "C and A by block rows over %d ranks, B halo rows exchanged (NCCL all_to_all)" % n_gpus,
            "l2": "inputs (A+B tiles, 5.8 GB) far exceed the 126 MB L2; no flush needed"}


def native_arm(args):
    import torch
    import torch.distributed as dist
    import hierarchical_block_sparse_lib_b200 as hb
    from hierarchical_block_sparse_lib_b200 import generators as G
    H = hb.HierarchicalBlockSparseMatrix

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        from hierarchical_block_sparse_lib_b200 import sharded
        return sharded.bench_main(args, WORKLOAD, workload_config, fp64_peak, ClockSampler)

    torch.cuda.set_device(local_rank)
    hb.init(local_rank)
    w = WORKLOAD
    n, b, lam, tau = w["n"], w["b"], w["lam"], w["tau"]
    W = G.decay_width(lam, w["eps"])
    A = H(np.float64, b); A.generate_decay(n, lam, W, w["seeds"][0]); A.update_internal_info()
    B = H(np.float64, b); B.generate_decay(n, lam, W, w["seeds"][1]); B.update_internal_info()

    def step():
        C = H(np.float64)
        nm, nr = H.spamm(A, False, B, False, C, tau, True)
        st = hb.stage_times()
        return C, nm, nr, st

    for _ in range(args.warmup):
        C, nm, nr, st = step()
        del C
    stream = torch.cuda.ExternalStream(_capi.lib().hbsm_stream())
    sampler = ClockSampler(local_rank); sampler.start()
    torch.cuda.synchronize()
    l0 = hb.kernel_launch_count()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    gemm_ms, task_ms = [], []
    ev0.record(stream)
    for _ in range(args.steps):
        C, nm, nr, st = step()
        gemm_ms.append(st["gemm_ms"]); task_ms.append(st["tasklist_ms"])
        del C
    ev1.record(stream)
    torch.cuda.synchronize()
    launches = hb.kernel_launch_count() - l0
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1) / args.steps
    flops = 2.0 * b ** 3 * nm
    value = flops / (ms * 1e-3) / 1e12
    peak, peak_src = fp64_peak()
    g_ms = float(np.mean(gemm_ms))
    achieved = flops / (g_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": "k_gemm_f64<64,64> (FP64 DMMA leaf GEMM)", "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak, "peak_source": peak_src,
                "algorithmic": "2*b^3 flops per leaf product x %d products per launch" % nm,
                "kernel_ms": g_ms, "share_of_step": g_ms / ms, "traffic": traffic_from_profile()}

    # ---- e2e: host buffers through the C ABI ----
    e2e = None if args.no_e2e else measure_e2e(hb, H, A, B, w, max(1, min(args.steps, 3)), torch)

    # ---- CPU baseline beside it (bounded sample) ----
    cpu = None
    if not args.no_cpu_baseline:
        try:
            tf, cms, info = run_reference_cpu(1, 1)
            cpu = {"value": tf, "unit": "TFLOP/s", "cores": info["cores"], "kind": info["kind"], "sample": info["sample"],
                   "ms_per_multiply_on_sample": cms}
        except Exception as ex:  # the checker is optional equipment; the GPU number stands without it
            cpu = {"value": None, "unit": "TFLOP/s", "cores": host_threads(), "kind": "unavailable", "sample": repr(ex)}

    line = {"metric": "spamm_fp64_leaf_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(1),
            "products_per_multiply": nm, "c_tiles": nr, "candidates": st["n_candidates"],
            "stage_ms": {"tasklist": float(np.mean(task_ms)), "gemm": g_ms},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
    print(json.dumps(line), flush=True)


def traffic_from_profile():
    p = os.path.join(ROOT, "profiles", "r01_gemm_f64_b64_traffic.json")
    try:
        return json.load(open(p))["dram_bytes_per_launch"]
    except Exception:
        return None


def measure_e2e(hb, H, A, B, w, steps, torch):
    """Same multiply through the C ABI with HOST buffers: pinned H2D of the A and B tiles, norm refresh, SpAMM, D2H
    of every C tile."""
    b, n, tau = w["b"], w["n"], w["tau"]

    def pinned_leaves(Mx):
        bi, bj, _, t = Mx.export_leaves(norms=False)
        pt = torch.empty(t.shape, dtype=torch.float64, pin_memory=True)
        pt.numpy()[...] = t
        return bi.astype(np.int32), bj.astype(np.int32), pt

    abi, abj, at = pinned_leaves(A)
    bbi, bbj, bt = pinned_leaves(B)
    h2d = at.numel() * 8 + bt.numel() * 8 + 4 * (len(abi) + len(abj) + len(bbi) + len(bbj))
    out = None
    times = []
    times_pipe = []
    d2h = 0
    nm = 0
    n_slabs = int(os.environ.get("HBSM_E2E_SLABS", "0"))
    import ctypes as Ct
    # schedule: two-call warm-up (learns the size of C), `steps` two-call passes, two pipelined warm-ups, `steps` pipelined
    # passes -- not interleaved, so that the stream-ordered memory pool is in steady state for both
    schedule = [False] * (steps + 1) + [True] * (steps + 2)
    untimed = {0, steps + 1, steps + 2}
    for i, pipelined in enumerate(schedule):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        A2 = H(np.float64, b); A2.resize(n, n)
        B2 = H(np.float64, b); B2.resize(n, n)
        C = H(np.float64)
        m = Ct.c_size_t(0)
        if pipelined:
            # ONE call: uploads, norm refresh, task lists, leaf GEMMs and downloads overlapped slab by slab
            nm, nr, cbi, cbj = H.product_from_host(A2, abi, abj, at.numpy(), False, B2, bbi, bbj, bt.numpy(), False, C, True, tau,
                                                   out.numpy(), n_slabs)
        else:
            A2.assign_tiles(abi, abj, at.numpy()); A2.update_internal_info()
            B2.assign_tiles(bbi, bbj, bt.numpy()); B2.update_internal_info()
            if out is None:      # warm-up pass: learn the size of C, allocate the pinned result buffer once
                nm, nr = H.spamm(A2, False, B2, False, C, tau, True)
                out = torch.empty((nr + nr // 8, b * b), dtype=torch.float64, pin_memory=True)
                _capi.check(_capi.lib().hbsm_export_leaves(C._h, nr, None, None, None, Ct.c_void_p(out.data_ptr()), Ct.byref(m)))
            else:                # SpAMM with the C tiles streaming to pinned host memory while the remaining leaf GEMMs run
                cnm = Ct.c_size_t(0); cnr = Ct.c_size_t(0)
                _capi.check(_capi.lib().hbsm_product_to_host(A2._h, 0, B2._h, 0, C._h, 1, float(tau), 1, Ct.c_void_p(out.data_ptr()),
                                                              out.shape[0], Ct.byref(cnm), Ct.byref(cnr)))
                nm, nr = cnm.value, cnr.value
            cbi = np.zeros(nr, np.int64); cbj = np.zeros(nr, np.int64)
            _capi.check(_capi.lib().hbsm_export_leaves(C._h, nr, cbi.ctypes.data_as(Ct.c_void_p), cbj.ctypes.data_as(Ct.c_void_p),
                                                        None, None, Ct.byref(m)))
        dt = time.perf_counter() - t0
        d2h = nr * b * b * 8 + 16 * nr
        del A2, B2, C
        if i not in untimed:
            (times_pipe if pipelined else times).append(dt)
    ms_two = 1e3 * float(np.mean(times))
    ms = 1e3 * float(np.mean(times_pipe))
    return {"value": 2.0 * b ** 3 * nm / (ms * 1e-3) / 1e12, "unit": "TFLOP/s", "ms_per_step": ms,
            "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "path": "hbsm_product_from_host: A, B tiles from pinned host memory, norm refresh, SpAMM task lists, leaf GEMMs and "
                    "the D2H of every C tile (+ coordinates) pipelined over block-row slabs of C on three streams",
            "ms_per_step_unpipelined": ms_two,
            "unpipelined_path": "hbsm_assign_tiles(A,B) + hbsm_update_norms + hbsm_product_to_host + hbsm_export_leaves(C keys)"}


def main():
    # stdout carries exactly ONE line (the JSON): everything else that writes to fd 1 (e.g. NCCL's version banner) is sent
    # to stderr; print() below writes through the saved descriptor
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-to-host measurement (used for the ncu launch list of the timed region)")
    ap.add_argument("--n", "--size", dest="n", type=int, default=None, help="override N (the judged run uses the default); use --size under torchrun")
    ap.add_argument("--lam", type=float, default=None)
    ap.add_argument("--leaf", type=int, default=None, help="override the leaf size (e.g. BASELINE config 4: --n 262144 --leaf 128)")
    args = ap.parse_args()
    if args.n:
        WORKLOAD["n"] = args.n
    if args.lam:
        WORKLOAD["lam"] = args.lam
    if args.leaf:
        WORKLOAD["b"] = args.leaf
    if args.impl == "reference":
        reference_arm(args)
    else:
        native_arm(args)


if __name__ == "__main__":
    main()
