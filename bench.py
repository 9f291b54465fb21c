#!/usr/bin/env python
"""bench.py — MC-dropout GA-MIL head throughput (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload config2|config3|config4]

Workload at the default (`config2`, BASELINE.json configs[1]): bags of N=1024 patches x 512-d
features, T=100 MC-dropout passes, 2 heads, shared attention.  A *step* is one pass of the hot
path (feature packing -> tcgen05 projection -> softmax rows -> Welford columns) over one packed
batch of `--bags-per-step` such bags (256 by default = 512 MB of fp32 features per GPU, larger
than the 126 MB L2, so no L2 flush is needed between timed steps).  N > 1: every rank owns its
own batch (bags are independent units: weak scaling, no data-path collective).

Prints ONE JSON line (see the task contract): `value` = bags/s with inputs resident in HBM,
`e2e` = the same metric through the public Python API with pinned-host features copied in and
results copied out inside the timed region, `roofline` for the projection kernel (CUDA events
around it, live), `cpu_baseline` = the torch-CPU port of the reference head on this host.
`--impl reference` times that port alone (the reference itself is Python and stays in the
build container; oracle/torch_port.py is bit-identical to it for the same seed).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L, D, C = 512, 128, 2
METRIC = "mc_dropout_gamil_bags_per_sec_T100_N1024"


def flops_per_bag(N, T, S):
    """F = T*N*(S*262144 + 512 + 2048) + T*4*L  (2 per MAC; tanh/exp/RNG not counted)."""
    return T * N * (S * 262144 + 512 + 2048) + T * 4 * L


def make_state_dict(seed, shared=True):
    g = torch.Generator().manual_seed(seed)

    def lin(o, i, bias=True):
        b = 1.0 / (i ** 0.5)
        w = (torch.rand(o, i, generator=g) * 2 - 1) * b
        return w, ((torch.rand(o, generator=g) * 2 - 1) * b if bias else None)

    sd = {}
    if shared:
        for nm in ("attention_V", "attention_U"):
            sd[f"{nm}.0.weight"], sd[f"{nm}.0.bias"] = lin(D, L)
    else:
        for nm in ("attention_V", "attention_U"):
            for c in range(C):
                sd[f"{nm}.{c}.0.weight"], sd[f"{nm}.{c}.0.bias"] = lin(D, L)
    for c in range(C):
        sd[f"attention_weights.{c}.weight"], sd[f"attention_weights.{c}.bias"] = lin(1, D)
    for c in range(C):
        sd[f"classifiers.{c}.weight"], _ = lin(1, L, bias=False)
    return sd


class ClockSampler:
    """SM clock + throttle reasons during the timed region (NVML)."""
    REASONS = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
               0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
               0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.stop_flag, self.thread = [], set(), False, None
        self.max_mhz, self.power = None, []
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def _run(self):
        while not self.stop_flag:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.power.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    self.nv, "nvmlDeviceGetCurrentClocksEventReasons") else self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.02)

    def start(self):
        if self.nv is not None:
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "power_w_max": (max(self.power) if self.power else None), "reasons": sorted(self.reasons),
                "samples": len(s)}


def workload(name, bags_per_step, seed):
    """(lengths per bag, T) of one rank's batch."""
    if name == "config2":
        return [1024] * bags_per_step, 100
    if name == "config3":   # 256 bags, N ~ U{200..3000}, T=50 (bag-sharded by LPT in run_ours)
        return [int(v) for v in np.random.default_rng(seed).integers(200, 3001, 256)], 50
    if name == "config4":   # one bag N=16384, T=1000 (MC samples sharded over the ranks)
        return [16384], 1000
    raise SystemExit(f"unknown workload {name}")


# =============================================================================================== reference arm
def cpu_head_bags_per_s(n_bags, N, T, shared, threads, warmup=1):
    from oracle import torch_port as TP
    torch.set_num_threads(threads)
    sd = make_state_dict(0, shared)
    g = torch.Generator().manual_seed(1)
    Hs = [torch.relu(torch.randn(N, L, generator=g)) for _ in range(min(n_bags, 4))]
    for i in range(warmup):
        TP.mc_head_torch(sd, Hs[i % len(Hs)], T, 0.1, 0.1)
    times = []
    for i in range(n_bags):
        t0 = time.perf_counter()
        TP.mc_head_torch(sd, Hs[i % len(Hs)], T, 0.1, 0.1)      # native torch dropout: the reference's true path
        times.append(time.perf_counter() - t0)
    return n_bags / sum(times), times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    lens, T = workload(args.workload, 1, args.seed)
    N = lens[0]
    cores = os.cpu_count() or 1
    T_run = T if args.workload != "config4" else 25     # the reference cannot materialise (1000,1,16384,512)
    t0 = time.perf_counter()
    bps, times = cpu_head_bags_per_s(args.steps, N, T_run, not args.separate, cores, warmup=args.warmup)
    if T_run != T:
        bps *= T_run / T
    ms = 1e3 * sum(times) / len(times) * (T / T_run)
    sample = f"{args.steps} bag(s) of N={N}, T={T_run}, one per step, torch-CPU port of model.py:280-316 with native dropout"
    line = {
        "impl": "reference", "metric": METRIC, "value": bps, "unit": "bags/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: N={N} patches x 512, T={T}, 2 heads, "
                               f"{'separate' if args.separate else 'shared'} attention; 1 bag per step on host cores"},
        "cpu_baseline": {"value": bps, "unit": "bags/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": bps, "unit": "bags/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


# =============================================================================================== our arm
def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
            os.environ["NCCL_DEBUG"] = "NONE"       # keep stdout to the ONE JSON line: at VERSION and WARN level NCCL
                                                    # prints its version banner there
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import mcmil_b200 as mm
    from mcmil_b200 import _lib
    from mcmil_b200 import distributed as MD
    lib = _lib.load()

    shared = not args.separate
    S = 1 if shared else C
    w = mm.HeadWeights(make_state_dict(0, shared), dev)
    all_lens, T = workload(args.workload, args.bags_per_step, args.seed)
    T_job = T
    t_offset, bag_ids = 0, None
    if args.workload == "config3" and world > 1:      # strong scaling: LPT bag sharding, no collective
        mine = MD.lpt_assign(all_lens, world)[rank]
        lens, bag_ids = [all_lens[i] for i in mine], mine
    elif args.workload == "config4" and world > 1:    # strong scaling: MC-sample sharding + one allreduce
        t_offset, T = MD.mc_shard(T, rank, world)
        lens = all_lens
    else:
        lens = all_lens
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    R = int(cu[-1])
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    H = torch.relu(torch.randn(R, L, generator=g, device=dev))       # synthetic ResNet-like features (>= 0)
    n_bags = len(lens)
    flops_step = sum(flops_per_bag(n, T, S) for n in lens)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    merged_T = [T]

    def step(i, rounds=None):
        r = mm.mc_head(w, H, T, seed=i, cu_seqlens=cu, bag_ids=bag_ids, t_offset=t_offset,
                       philox_rounds=rounds or args.philox_rounds)
        if args.workload == "config4" and world > 1:
            _, _, merged_T[0] = MD.allreduce_welford([r.attn_mean, r.prob_mean], [r.attn_m2, r.prob_m2], T,
                                                     total_count=T_job)
        return r

    for i in range(max(args.warmup, 3)):
        res = step(i)
    launches_per_step = res.launches
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    _lib.check(lib.mcmil_profile_begin(args.steps), "mcmil_profile_begin")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(100 + i)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    prof_ms, prof_k = ctypes.c_double(0), ctypes.c_int(0)
    _lib.check(lib.mcmil_profile_end(ctypes.byref(prof_ms), ctypes.byref(prof_k)), "mcmil_profile_end")
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    strong = args.workload in ("config3", "config4") and world > 1
    job_bags = len(all_lens) if strong else n_bags * world
    value = job_bags * args.steps / (ms_total / 1e3)

    # ---- roofline of the dominant kernel (tcgen05 projection), CUDA events around its launches
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:  # noqa: BLE001
        pass
    peak = float(peaks.get("bf16_tflops", 1590.0))
    peak_sust = float(peaks.get("bf16_tflops_sustained", 1400.0))
    k_ms = prof_ms.value / max(prof_k.value, 1)
    achieved = (flops_step / S) / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
    # Peak: MEASURED_PEAKS.json holds a burst figure (one cuBLAS bf16 GEMM timed alone) and a sustained one (GEMMs back
    # to back for seconds).  The projection kernel is timed inside a long step (K steps of ~6 ms back to back, each
    # ~96 % this kernel), so the sustained figure is the roofline; the fraction of the burst figure is reported too.
    roofline = {"bound": "tensor", "kernel": "proj_tc_kernel", "achieved": achieved, "peak": peak_sust, "unit": "TFLOP/s",
                "frac": achieved / peak_sust,
                "peak_source": ("measured (MEASURED_PEAKS.json bf16 sustained: kernel timed inside a long step)" if peaks
                                else "fallback"),
                "peak_burst": peak, "frac_of_burst": achieved / peak, "kernel_ms": k_ms, "kernel_launches": prof_k.value,
                "kernel_share_of_step": prof_ms.value / ms_total if ms_total > 0 else None, "traffic": None}
    # DRAM traffic per launch of that kernel, from the committed ncu capture of this very configuration
    # (profiles/r1_traffic.json; ncu is never attached to a timed run)
    if args.workload == "config2" and args.bags_per_step == 256 and shared and args.philox_rounds == 10:
        try:
            with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
                tr = json.load(f)["config2_shared"]
            roofline["traffic"] = tr["dram_read_bytes"] + tr["dram_write_bytes"]
            roofline["traffic_unit"] = "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/r1_traffic.json)"
            roofline["algorithmic_bytes"] = sum(tr["algorithmic_bytes"].values())
        except Exception:  # noqa: BLE001
            pass

    # ---- the same step with Philox4x32-7 masks (optional fast mode; not the headline)
    philox7 = None
    if args.philox_rounds == 10 and not args.no_extras:
        for i in range(3):
            step(i, 7)
        barrier()
        _lib.check(lib.mcmil_profile_begin(args.steps), "mcmil_profile_begin")
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for i in range(args.steps):
            step(300 + i, 7)
        f1.record()
        barrier()
        p7_ms, p7_k = ctypes.c_double(0), ctypes.c_int(0)
        _lib.check(lib.mcmil_profile_end(ctypes.byref(p7_ms), ctypes.byref(p7_k)), "mcmil_profile_end")
        ms7 = f0.elapsed_time(f1)
        k7 = p7_ms.value / max(p7_k.value, 1)
        philox7 = {"bags_per_s_this_rank": n_bags * args.steps / (ms7 / 1e3),
                   "kernel_ms": k7, "roofline_frac": (flops_step / S) / (k7 * 1e-3) / 1e12 / peak_sust if k7 > 0 else None}

    # ---- e2e: pinned-host features in, results out, through the public API, inside the timed region
    e2e, e2e_f16 = None, None

    def measure_e2e(dtype):
        chunk_bags = max(1, min(n_bags, args.e2e_chunk))
        H_host = torch.empty((R, L), dtype=dtype).pin_memory()
        H_host.copy_(H.to(dtype).cpu())
        streams = [torch.cuda.Stream(dev) for _ in range(2)]
        n_chunks = (n_bags + chunk_bags - 1) // chunk_bags
        max_rows = max(int(cu[min(n_bags, (k + 1) * chunk_bags)] - cu[k * chunk_bags]) for k in range(n_chunks))
        dbuf = [torch.empty((max_rows, L), dtype=dtype, device=dev) for _ in range(2)]
        # results land in flat pinned buffers (one per chunk): every D2H copy is a single contiguous
        # cudaMemcpyAsync (a strided pinned destination makes torch stage + synchronise, which
        # serialises the whole pipeline)
        outY = [None] * n_chunks
        outP = [None] * n_chunks
        outA = [None] * n_chunks
        for k in range(n_chunks):
            b0, b1 = k * chunk_bags, min(n_bags, (k + 1) * chunk_bags)
            rows = int(cu[b1] - cu[b0])
            outY[k] = torch.empty((b1 - b0) * T * C, dtype=torch.float32).pin_memory()
            outP[k] = torch.empty(2 * (b1 - b0) * C, dtype=torch.float32).pin_memory()
            outA[k] = torch.empty(2 * C * rows, dtype=torch.float32).pin_memory()
        h2d = R * L * H_host.element_size()
        d2h = sum(t.numel() for t in outY + outP + outA) * 4

        def e2e_step(i):
            for k in range(n_chunks):
                b0, b1 = k * chunk_bags, min(n_bags, (k + 1) * chunk_bags)
                r0, r1 = int(cu[b0]), int(cu[b1])
                s = streams[k % 2]
                with torch.cuda.stream(s):
                    hb = dbuf[k % 2][: r1 - r0]
                    hb.copy_(H_host[r0:r1], non_blocking=True)
                    r = mm.mc_head(w, hb, T, seed=i, cu_seqlens=cu[b0:b1 + 1] - cu[b0],
                                   bag_ids=None if bag_ids is None else bag_ids[b0:b1], bag_offset=b0,
                                   t_offset=t_offset, philox_rounds=args.philox_rounds)
                    nb_c, na_c = (b1 - b0) * C, C * (r1 - r0)
                    outY[k].copy_(r.Y.view(-1), non_blocking=True)
                    outP[k][:nb_c].copy_(r.prob_mean.view(-1), non_blocking=True)
                    outP[k][nb_c:].copy_(r.prob_m2.view(-1), non_blocking=True)
                    outA[k][:na_c].copy_(r.attn_mean.view(-1), non_blocking=True)
                    outA[k][na_c:].copy_(r.attn_m2.view(-1), non_blocking=True)
            for s in streams:
                s.synchronize()

        for i in range(3):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            e2e_step(200 + i)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return {"value": job_bags * args.steps / dt, "unit": "bags/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "chunk_bags": chunk_bags, "streams": 2,
                "api": "mcmil_b200.mc_head on pinned-host %s features (copy in, compute, copy out, pipelined over 2 streams)"
                       % ("float32" if dtype == torch.float32 else "float16")}

    if not args.no_e2e:
        e2e = measure_e2e(torch.float32)                 # the reference's feature dtype: the e2e number of record
        if not args.no_extras:
            e2e_f16 = measure_e2e(torch.float16)         # features handed over in half precision (same results)

    # ---- single-bag call latency / back-to-back throughput (the reference's bs=1 usage)
    single = None
    if args.workload == "config2" and rank == 0 and not args.no_extras:
        nb = min(n_bags, 128)
        for i in range(10):
            mm.mc_head(w, H[(i % nb) * 1024:(i % nb + 1) * 1024], T, seed=i)
        torch.cuda.synchronize(dev)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 200
        s0.record()
        for i in range(reps):
            mm.mc_head(w, H[(i % nb) * 1024:(i % nb + 1) * 1024], T, seed=i)
        s1.record()
        torch.cuda.synchronize(dev)
        per = s0.elapsed_time(s1) / reps
        single = {"us_per_bag_back_to_back": per * 1e3, "bags_per_s": 1e3 / per, "calls": reps}
        runner = mm.MCHeadRunner(w, 1024, T)           # same call with plan / outputs created once
        for i in range(10):
            runner.run(H[(i % nb) * 1024:(i % nb + 1) * 1024], seed=i)
        torch.cuda.synchronize(dev)
        s0.record()
        for i in range(reps):
            runner.run(H[(i % nb) * 1024:(i % nb + 1) * 1024], seed=i)
        s1.record()
        torch.cuda.synchronize(dev)
        per = s0.elapsed_time(s1) / reps
        single["runner_us_per_bag"] = per * 1e3
        single["runner_bags_per_s"] = 1e3 / per

    # ---- CPU baseline on this host (rank 0, N=1 only): bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        Ncpu, Tcpu = (1024, 100) if args.workload == "config2" else (lens[0], min(T, 25))
        bps, times = cpu_head_bags_per_s(args.cpu_bags, Ncpu, Tcpu, shared, cores)
        cpu = {"value": bps * (Tcpu / (100 if args.workload == "config2" else T)), "unit": "bags/s", "cores": cores,
               "kind": "port",
               "sample": f"{args.cpu_bags} bag(s) of N={Ncpu}, T={Tcpu}: torch-CPU port of the reference head "
                         f"(oracle/torch_port.py, native dropout), {cores} threads"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "bags/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f16xf16->f32",
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: {len(all_lens) if strong else n_bags} bag(s)"
                                   f"{'' if strong else ' per GPU'} per step, N={'1024' if args.workload == 'config2' else 'var'} "
                                   f"patches x 512 features, T={merged_T[0] if strong else T} MC passes, {C} heads, "
                                   f"{'shared' if shared else 'separate'} attention, p_f=p_a=0.1, in-kernel Philox masks",
                       "bags_per_step_per_gpu": n_bags, "rows_per_step_per_gpu": R,
                       "l2_policy": "inputs larger than L2 (%.0f MB of features per step per GPU), no flush" % (R * L * 4 / 1e6),
                       "parallelism": "bags sharded over ranks, no collective" if args.workload != "config4"
                       else "MC samples sharded over ranks, one NCCL allreduce of Welford partials"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_f16_features": e2e_f16, "single_bag": single,
            "philox_rounds": args.philox_rounds, "philox7_mode": philox7,
            "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
            "clocks": clocks, "flops_per_step_per_gpu": flops_step,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=["config2", "config3", "config4"])
    ap.add_argument("--bags-per-step", type=int, default=256)
    ap.add_argument("--separate", action="store_true", help="shared_attention=False (config.yml default)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--e2e-chunk", type=int, default=32)
    ap.add_argument("--cpu-bags", type=int, default=8)
    ap.add_argument("--philox-rounds", type=int, default=10, choices=[7, 10])
    ap.add_argument("--no-extras", action="store_true", help="skip the Philox-7 and single-bag extras")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
