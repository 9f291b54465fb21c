#!/usr/bin/env python
"""bench.py — MC-dropout GA-MIL head throughput (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload config2|config3|config4]

Workload at the default (`config2`, BASELINE.json configs[1]): bags of N=1024 patches x 512-d
features, T=100 MC-dropout passes, 2 heads, shared attention.  A *step* is one pass of the hot
path (tcgen05 projection -> softmax rows -> Welford columns) over one packed batch of
`--bags-per-step` such bags (256 by default = 512 MB of fp32 features per GPU, larger than the
126 MB L2, so no L2 flush is needed between timed steps).  N > 1: every rank owns its own batch
(bags are independent units: weak scaling, no data-path collective).

Prints ONE JSON line (see the task contract): `value` = bags/s with inputs resident in HBM,
`e2e` = the same metric through the public Python API with pinned-host features copied in and
results copied out inside the timed region, `roofline` for the projection kernel (CUDA events
around it, live), `cpu_baseline` = the reference's CPU path on this host.  Extras of the same line
(all measured in this run, every N): `single_bag` (the reference's bs == 1 call pattern),
`separate` (shared_attention=False, the reference's YAML default), `config3` (256 ragged bags,
LPT bag sharding, strong scaling, no collective), `config4` (one bag of 16384 patches, T=1000, MC
samples sharded over the ranks + ONE NCCL all-reduce of the Welford partials, with an on-device
check of the merged statistics against a single-rank run of the same global samples), `config5`
(N=1 only: image -> tiling -> torch ResNet-18 -> MC head -> attention-map statistics, stage times).

`--impl reference` times the reference's own CPU implementation: `baseline/_ref/model.py` (the
unmodified reference module, placed there by `__graft_entry__.build()` when `/root/reference`
exists) or `/root/reference/model.py`, else the torch-CPU port `oracle/torch_port.py`
(bit-identical to the reference for the same seed, tests/golden/make_golden.py).
"""
from __future__ import annotations

import argparse
import ctypes
import importlib.util
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
def load_reference_module():
    """The reference's own model.py: baseline/_ref (travels to the GPU box) or /root/reference (build container)."""
    for path in (os.path.join(ROOT, "baseline", "_ref", "model.py"), "/root/reference/model.py"):
        if os.path.exists(path):
            try:
                spec = importlib.util.spec_from_file_location("_mcmil_reference_model", path)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                return mod, path
            except Exception as e:  # noqa: BLE001  (e.g. torchvision missing): fall back to the port
                print(f"bench: could not import {path}: {e!r}", file=sys.stderr)
    return None, None


class CpuHead:
    """One bag through the reference's CPU path: the unmodified reference module when it is present
    (kind "reference": model.py:256-328 with `feature_extractor = nn.Flatten()` so the features go in
    directly, as SURVEY §8c), else the torch port (kind "port")."""

    def __init__(self, shared, threads):
        torch.set_num_threads(threads)
        self.sd = make_state_dict(0, shared)
        self.mod, self.path = load_reference_module()
        self.kind = "reference" if self.mod is not None else "port"
        if self.mod is not None:
            m = self.mod.MultiHeadGatedAttentionMIL(pretrained=False, shared_attention=shared)
            m.feature_extractor = torch.nn.Flatten()
            missing, unexpected = m.load_state_dict(self.sd, strict=False)
            assert not unexpected and all(k.startswith("feature_extractor") for k in missing), (missing, unexpected)
            self.model = m
        else:
            from oracle import torch_port as TP
            self.TP = TP

    def run(self, H, T):
        if self.mod is not None:
            return self.model.mc_inference(H.view(1, H.shape[0], L, 1, 1), N=T, device="cpu")
        return self.TP.mc_head_torch(self.sd, H, T, 0.1, 0.1)        # native torch dropout: the reference's true path

    def describe(self):
        return ("unmodified reference module %s (mc_inference, device='cpu', native dropout)" % os.path.relpath(self.path, ROOT)
                if self.mod is not None else "torch-CPU port of model.py:280-316 (oracle/torch_port.py, native dropout)")


def cpu_head_bags_per_s(n_bags, N, T, shared, threads, warmup=1):
    head = CpuHead(shared, threads)
    g = torch.Generator().manual_seed(1)
    Hs = [torch.relu(torch.randn(N, L, generator=g)) for _ in range(min(n_bags, 4))]
    for i in range(warmup):
        head.run(Hs[i % len(Hs)], T)
    times = []
    for i in range(n_bags):
        t0 = time.perf_counter()
        head.run(Hs[i % len(Hs)], T)
        times.append(time.perf_counter() - t0)
    return n_bags / sum(times), times, head


def config_of(name, n_bags, rows_per_gpu, T, C, shared, strong):
    """The `config` object of the JSON line: both arms print the same one (the reference arm measures the same
    metric on a bounded sample of this workload)."""
    return {"workload": f"{name}: {n_bags} bag(s){'' if strong else ' per GPU'} per step, "
                        f"N={'1024' if name == 'config2' else 'var'} patches x 512 features, T={T} MC passes, {C} heads, "
                        f"{'shared' if shared else 'separate'} attention, p_f=p_a=0.1",
            "bags_per_step_per_gpu": n_bags, "rows_per_step_per_gpu": rows_per_gpu,
            "l2_policy": "inputs larger than L2 (%.0f MB of features per step per GPU), no flush" % (rows_per_gpu * L * 4 / 1e6),
            "parallelism": "bags sharded over ranks, no collective" if name != "config4"
            else "MC samples sharded over ranks, one NCCL allreduce of Welford partials"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    lens, T = workload(args.workload, 1, args.seed)
    N = lens[0]
    cores = os.cpu_count() or 1
    # the line carries OUR arm's config (same metric on the same workload); what one step of this arm runs is a
    # bounded sample of it, described in cpu_baseline.sample
    all_lens, _ = workload(args.workload, args.bags_per_step, args.seed)
    strong = args.workload in ("config3", "config4") and args.gpus > 1
    ours_config = config_of(args.workload, len(all_lens), sum(all_lens), T, 2, not args.separate, strong)
    T_run = T if args.workload != "config4" else 25     # the reference cannot materialise (1000,1,16384,512)
    t0 = time.perf_counter()
    bps, times, head = cpu_head_bags_per_s(args.steps, N, T_run, not args.separate, cores, warmup=args.warmup)
    if T_run != T:
        bps *= T_run / T
    ms = 1e3 * sum(times) / len(times) * (T / T_run)
    sample = f"{args.steps} bag(s) of N={N}, T={T_run}, one per step: {head.describe()}, {cores} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": bps, "unit": "bags/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": ours_config,
        "cpu_baseline": {"value": bps, "unit": "bags/s", "cores": cores, "kind": head.kind, "sample": sample},
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

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:  # noqa: BLE001
        pass
    peak_burst = float(peaks.get("bf16_tflops", 1590.0))
    peak_sust = float(peaks.get("bf16_tflops_sustained", 1400.0))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(step_fn, steps, warm, profile=False):
        """`warm` untimed + `steps` timed calls of step_fn(i); CUDA events, barrier + synchronize on both sides,
        max over ranks.  profile: also the summed device time of the projection launches (library events)."""
        for i in range(warm):
            step_fn(i)
        barrier()
        if profile:
            _lib.check(lib.mcmil_profile_begin(steps * 2), "mcmil_profile_begin")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(steps):
            step_fn(1000 + i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        prof = None
        if profile:
            pm, pk = ctypes.c_double(0), ctypes.c_int(0)
            _lib.check(lib.mcmil_profile_end(ctypes.byref(pm), ctypes.byref(pk)), "mcmil_profile_end")
            prof = (pm.value, pk.value)
        return max_over_ranks(ms), prof

    # ------------------------------------------------------------------------------------------ main leg
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
    merged_T = [T]
    last = [None]

    def step(i, rounds=None):
        r = mm.mc_head(w, H, T, seed=i, cu_seqlens=cu, bag_ids=bag_ids, t_offset=t_offset,
                       philox_rounds=rounds or args.philox_rounds)
        if args.workload == "config4" and world > 1:
            _, _, merged_T[0] = MD.allreduce_welford([r.attn_mean, r.prob_mean], [r.attn_m2, r.prob_m2], T,
                                                     total_count=T_job)
        last[0] = r
        return r

    sampler = ClockSampler(local_rank)
    warm = max(args.warmup, 3)
    for i in range(warm):          # warm-up outside the clock sampling
        step(i)
    barrier()
    sampler.start()
    ms_total, (prof_ms, prof_k) = timed(step, args.steps, 0, profile=True)
    clocks = sampler.stop()
    launches_per_step = last[0].launches
    strong = args.workload in ("config3", "config4") and world > 1
    job_bags = len(all_lens) if strong else n_bags * world
    value = job_bags * args.steps / (ms_total / 1e3)

    # ---- roofline of the dominant kernel (tcgen05 projection), CUDA events around its launches.
    # Peak: MEASURED_PEAKS.json holds a burst figure (one cuBLAS bf16 GEMM timed alone) and a sustained one (GEMMs back
    # to back for seconds, power-capped).  The contract assigns the sustained figure to kernels timed inside a long
    # step; a timed region shorter than one second is not long, so the burst figure is the denominator there.
    k_ms = prof_ms / max(prof_k, 1)
    achieved = (flops_step / S) / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
    long_region = ms_total >= 1000.0
    peak = peak_sust if long_region else peak_burst
    roofline = {"bound": "tensor", "kernel": "proj_tc_kernel", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak,
                "peak_source": ("%s (MEASURED_PEAKS.json bf16 %s: timed region %.2f s)" % (
                    "measured" if peaks else "fallback", "sustained" if long_region else "burst", ms_total / 1e3)),
                "peak_burst": peak_burst, "frac_of_burst": achieved / peak_burst,
                "peak_sustained": peak_sust, "frac_of_sustained": achieved / peak_sust,
                "kernel_ms": k_ms, "kernel_launches": prof_k,
                "kernel_share_of_step": prof_ms / ms_total if ms_total > 0 else None,
                "algorithmic_flops_per_launch": flops_step / S, "traffic": None}
    # DRAM traffic per launch of that kernel: only from an ncu capture of THIS round's binary at this very
    # configuration (profiles/r2_traffic.json; ncu is never attached to a timed run); otherwise null.
    if args.workload == "config2" and args.bags_per_step == 256 and shared and args.philox_rounds == 10 and world == 1:
        try:
            with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
                tr = json.load(f)["config2_shared"]
            roofline["traffic"] = tr["dram_read_bytes"] + tr["dram_write_bytes"]
            roofline["traffic_unit"] = "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/r2_traffic.json)"
            roofline["algorithmic_bytes"] = sum(tr["algorithmic_bytes"].values())
        except Exception:  # noqa: BLE001
            pass

    extras_on = not args.no_extras and args.workload == "config2"

    # ---- the same step with Philox4x32-7 masks (optional fast mode; not the headline)
    philox7 = None
    if args.philox_rounds == 10 and extras_on and world == 1:
        ms7, (p7_ms, p7_k) = timed(lambda i: step(i, 7), args.steps, 3, profile=True)
        k7 = p7_ms / max(p7_k, 1)
        philox7 = {"bags_per_s_this_rank": n_bags * args.steps / (ms7 / 1e3),
                   "kernel_ms": k7, "roofline_frac": (flops_step / S) / (k7 * 1e-3) / 1e12 / peak if k7 > 0 else None}

    # ---- separate attention (shared_attention=False, the reference's config.yml default): same batch, S = 2
    separate = None
    if extras_on and shared:
        w_sep = mm.HeadWeights(make_state_dict(0, False), dev)

        def sep_step(i):
            return mm.mc_head(w_sep, H, T, seed=i, cu_seqlens=cu, philox_rounds=args.philox_rounds)

        steps_s = max(3, args.steps // 2)
        ms_s, (ps_ms, ps_k) = timed(sep_step, steps_s, 3, profile=True)
        fl = sum(flops_per_bag(n, T, C) for n in lens)          # both heads
        ach = fl * steps_s / (ps_ms * 1e-3) / 1e12 if ps_ms > 0 else 0.0
        separate = {"bags_per_s": n_bags * world * steps_s / (ms_s / 1e3), "ms_per_step": ms_s / steps_s,
                    "steps": steps_s, "projection_ms_per_step": ps_ms / steps_s,
                    "projection_launches_per_step": ps_k / steps_s,
                    "achieved_tflops": ach, "frac": ach / peak,
                    "flop_rate_vs_shared": (ach / achieved) if achieved > 0 else None}
        del w_sep

    # ---- e2e: pinned-host features in, results out, through the public API, inside the timed region
    e2e, e2e_f16 = None, None

    def measure_e2e(dtype):
        chunk_bags = max(1, min(n_bags, args.e2e_chunk))
        n_streams = args.e2e_streams
        H_host = torch.empty((R, L), dtype=dtype).pin_memory()
        H_host.copy_(H.to(dtype).cpu())
        streams = [torch.cuda.Stream(dev) for _ in range(n_streams)]
        n_chunks = (n_bags + chunk_bags - 1) // chunk_bags
        max_rows = max(int(cu[min(n_bags, (k + 1) * chunk_bags)] - cu[k * chunk_bags]) for k in range(n_chunks))
        dbuf = [torch.empty((max_rows, L), dtype=dtype, device=dev) for _ in range(n_streams)]
        # results land in flat pinned buffers (one per chunk): every D2H copy is a single contiguous
        # cudaMemcpyAsync (a strided pinned destination makes torch stage + synchronise, which
        # serialises the whole pipeline)
        # two sets (step parity): the host reads step i's results while step i+1 is already in flight
        outY = [[None] * n_chunks for _ in range(2)]
        outP = [[None] * n_chunks for _ in range(2)]
        outA = [[None] * n_chunks for _ in range(2)]
        for par in range(2):
            for k in range(n_chunks):
                b0, b1 = k * chunk_bags, min(n_bags, (k + 1) * chunk_bags)
                rows = int(cu[b1] - cu[b0])
                outY[par][k] = torch.empty((b1 - b0) * T * C, dtype=torch.float32).pin_memory()
                outP[par][k] = torch.empty(2 * (b1 - b0) * C, dtype=torch.float32).pin_memory()
                outA[par][k] = torch.empty(2 * C * rows, dtype=torch.float32).pin_memory()
        h2d = R * L * H_host.element_size()
        d2h = sum(t.numel() for t in outY[0] + outP[0] + outA[0]) * 4

        def enqueue(i):
            """Copy in, compute, copy out for every chunk of step i; returns the events that mark its results."""
            par = i & 1
            done = []
            for k in range(n_chunks):
                b0, b1 = k * chunk_bags, min(n_bags, (k + 1) * chunk_bags)
                r0, r1 = int(cu[b0]), int(cu[b1])
                s = streams[k % n_streams]
                with torch.cuda.stream(s):
                    hb = dbuf[k % n_streams][: r1 - r0]
                    hb.copy_(H_host[r0:r1], non_blocking=True)
                    r = mm.mc_head(w, hb, T, seed=i, cu_seqlens=cu[b0:b1 + 1] - cu[b0],
                                   bag_ids=None if bag_ids is None else bag_ids[b0:b1], bag_offset=b0,
                                   t_offset=t_offset, philox_rounds=args.philox_rounds)
                    nb_c, na_c = (b1 - b0) * C, C * (r1 - r0)
                    outY[par][k].copy_(r.Y.view(-1), non_blocking=True)
                    outP[par][k][:nb_c].copy_(r.prob_mean.view(-1), non_blocking=True)
                    outP[par][k][nb_c:].copy_(r.prob_m2.view(-1), non_blocking=True)
                    outA[par][k][:na_c].copy_(r.attn_mean.view(-1), non_blocking=True)
                    outA[par][k][na_c:].copy_(r.attn_m2.view(-1), non_blocking=True)
                    if k >= n_chunks - n_streams:          # the last chunk of every stream
                        ev = torch.cuda.Event()
                        ev.record(s)
                        done.append(ev)
            return done

        def consume(i, done):
            """The host side of step i: wait for its results and read them (one value per chunk stands for the reader)."""
            for ev in done:
                ev.synchronize()
            par = i & 1
            return sum(float(outP[par][k][0]) for k in range(n_chunks))

        def run_steps(first, count, overlap):
            # overlap: step i+1 is enqueued before step i's results are read (two result sets), so the host link
            # never idles between steps; otherwise every step is drained before the next one starts
            prev = None
            for i in range(first, first + count):
                done = enqueue(i)
                if not overlap:
                    consume(i, done)
                    continue
                if prev is not None:
                    consume(*prev)
                prev = (i, done)
            if prev is not None:
                consume(*prev)

        results = {}
        for overlap in (False, True):
            run_steps(0, 3, overlap)
            torch.cuda.synchronize(dev)
            barrier()
            t0 = time.perf_counter()
            run_steps(200, args.steps, overlap)
            torch.cuda.synchronize(dev)
            results[overlap] = max_over_ranks(time.perf_counter() - t0)
        dt = results[True]
        return {"value": job_bags * args.steps / dt, "unit": "bags/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "chunk_bags": chunk_bags, "streams": n_streams,
                "h2d_gbs_per_gpu": h2d * args.steps / dt / 1e9,
                "value_step_drained": job_bags * args.steps / results[False],
                "api": "mcmil_b200.mc_head on pinned-host %s features (copy in, compute, copy out, pipelined over %d "
                       "streams; the host reads step i's results while step i+1 is in flight - value_step_drained: every "
                       "step drained before the next one starts)" % ("float32" if dtype == torch.float32 else "float16", n_streams)}

    if not args.no_e2e:
        e2e = measure_e2e(torch.float32)                 # the reference's feature dtype: the e2e number of record
        if extras_on:
            e2e_f16 = measure_e2e(torch.float16)         # features handed over in half precision (same results)
            # the copy alone: what the host link gives this rank while all ranks copy at once (the e2e limiter)
            hp = torch.empty((64 << 20) // 4, dtype=torch.float32).pin_memory()
            dp = torch.empty_like(hp, device=dev)
            for _ in range(2):
                dp.copy_(hp, non_blocking=True)
            best = 0.0
            for _ in range(3):                    # best of 3 x 16 copies of 64 MB, all ranks copying at the same time
                barrier()
                t0 = time.perf_counter()
                for _ in range(16):
                    dp.copy_(hp, non_blocking=True)
                torch.cuda.synchronize(dev)
                best = max(best, 16 * hp.numel() * 4 / max_over_ranks(time.perf_counter() - t0) / 1e9)
            e2e["h2d_copy_only_gbs_per_gpu"] = best
            del hp, dp

    # ---- single-bag call latency / back-to-back throughput (the reference's bs == 1 serving loop, infer.py:187-196)
    single = None
    if extras_on and rank == 0:
        nb = min(n_bags, 128)
        reps = 200
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def loop(fn):
            for i in range(20):
                fn(i)
            torch.cuda.synchronize(dev)
            s0.record()
            for i in range(reps):
                fn(i)
            s1.record()
            torch.cuda.synchronize(dev)
            return s0.elapsed_time(s1) / reps

        per = loop(lambda i: mm.mc_head(w, H[(i % nb) * 1024:(i % nb + 1) * 1024], T, seed=i))
        single = {"us_per_bag_back_to_back": per * 1e3, "bags_per_s": 1e3 / per, "calls": reps}
        runner = mm.MCHeadRunner(w, 1024, T)           # same call with plan / outputs created once
        per = loop(lambda i: runner.run(H[(i % nb) * 1024:(i % nb + 1) * 1024], seed=i))
        f1 = flops_per_bag(1024, T, S)
        single.update({"runner_us_per_bag": per * 1e3, "runner_bags_per_s": 1e3 / per,
                       "launches_per_bag": int(lib.mcmil_last_launch_count()),
                       "achieved_tflops": f1 / (per * 1e-3) / 1e12, "frac_of_burst": f1 / (per * 1e-3) / 1e12 / peak_burst,
                       "roofline_us": f1 / (peak_burst * 1e12) * 1e6})
        # one isolated call (device time of a single call with an idle GPU before it)
        iso = []
        for i in range(10):
            torch.cuda.synchronize(dev)
            s0.record()
            runner.run(H[:1024], seed=i)
            s1.record()
            torch.cuda.synchronize(dev)
            iso.append(s0.elapsed_time(s1) * 1e3)
        single["isolated_call_us_median"] = sorted(iso)[len(iso) // 2]
        # throughput mode: the same single-bag calls round-robin over k private streams, each projection kernel on
        # 1/k of the SMs, so the fixed per-kernel cost of one call overlaps with the other bags (MCHeadRunner docstring)
        tp = {}
        for k in (4, 8):
            rk = mm.MCHeadRunner(w, 1024, T, n_streams=k)
            # (sync_input=False: the features were written long before and are not touched again, so the private
            # streams need not wait for the caller's stream on every call)
            for i in range(40):
                rk.run(H[(i % nb) * 1024:(i % nb + 1) * 1024], seed=i, sync_input=False)
            rk.synchronize()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for i in range(2 * reps):
                rk.run(H[(i % nb) * 1024:(i % nb + 1) * 1024], seed=i, sync_input=False)
            t_issue = time.perf_counter() - t0
            rk.synchronize()
            dt = time.perf_counter() - t0
            tp[f"streams_{k}"] = {"us_per_bag": dt / (2 * reps) * 1e6, "bags_per_s": 2 * reps / dt,
                                  "host_issue_us_per_call": t_issue / (2 * reps) * 1e6,
                                  "frac_of_burst": f1 / (dt / (2 * reps)) / 1e12 / peak_burst}
            # the same calls with no Python in the loop (the host loop above is host bound on slow hosts): every
            # private stream replays a CUDA graph of 16 of its calls
            graphs = []
            for slot in rk.slots:
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.stream(slot.stream):
                    with torch.cuda.graph(gr, stream=slot.stream):
                        for i in range(16):
                            _lib.check(rk.lib.mcmil_head_forward(rk.w._h, slot.plan._h, H[(i % nb) * 1024:].data_ptr(), 0, 0, i,
                                                                 *slot.tail, slot.stream.cuda_stream), "mcmil_head_forward")
                graphs.append((gr, slot.stream))
            for gr, st_ in graphs:
                with torch.cuda.stream(st_):
                    gr.replay()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for _ in range(5):
                for gr, st_ in graphs:
                    with torch.cuda.stream(st_):
                        gr.replay()
            torch.cuda.synchronize(dev)
            dtg = (time.perf_counter() - t0) / (5 * 16 * k)
            tp[f"streams_{k}"].update({"graph_replay_us_per_bag": dtg * 1e6, "graph_replay_bags_per_s": 1.0 / dtg})
            del graphs, rk
        single["throughput_mode"] = tp
        # the back-to-back rate of single-bag calls a serving loop reaches with 8 calls in flight (Python host loop)
        single["bags_per_s_8_streams"] = tp["streams_8"]["bags_per_s"]
        single["us_per_bag_8_streams"] = tp["streams_8"]["us_per_bag"]

    # ---- configs 3 and 4 of BASELINE.json, strong scaling over the ranks of this run
    def config3_leg():
        lens3, T3 = workload("config3", 0, args.seed)
        mine = MD.lpt_assign(lens3, world)[rank]
        my_lens = [lens3[i] for i in mine]
        cu3 = np.concatenate([[0], np.cumsum(my_lens)]).astype(np.int32)
        g3 = torch.Generator(device=dev).manual_seed(77 + rank)
        H3 = torch.relu(torch.randn(int(cu3[-1]), L, generator=g3, device=dev))
        steps3 = max(3, min(args.steps, 10))
        ms3, (p_ms, p_k) = timed(lambda i: mm.mc_head(w, H3, T3, seed=i, cu_seqlens=cu3, bag_ids=mine), steps3, 3,
                                 profile=True)
        fl_rank = sum(flops_per_bag(n, T3, S) for n in my_lens)
        tiles = [sum(-(-n // 128) for n in [lens3[i] for i in MD.lpt_assign(lens3, world)[r]]) for r in range(world)]
        return {"scaling": "strong", "bags": len(lens3), "T": T3, "n_gpus": world, "steps": steps3,
                "ms_per_step": ms3 / steps3, "bags_per_s": len(lens3) * steps3 / (ms3 / 1e3),
                "projection_ms_per_step_this_rank": p_ms / steps3,
                "achieved_tflops_this_rank": fl_rank / (p_ms / steps3 * 1e-3) / 1e12 if p_ms > 0 else None,
                "tiles_per_rank_min_max": [min(tiles), max(tiles)], "collective": "none (bags are independent units)"}

    def config4_leg():
        lens4, T4 = workload("config4", 0, args.seed)
        N4 = lens4[0]
        t0_, Tl = MD.mc_shard(T4, rank, world)
        g4 = torch.Generator(device=dev).manual_seed(4242)           # H is replicated: same seed on every rank
        H4 = torch.relu(torch.randn(N4, L, generator=g4, device=dev))
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        split_ms = [0.0, 0.0]
        merged = [None]

        def step4(i):
            timing = i >= 1000
            if timing:
                ev[0].record()
            r = mm.mc_head(w, H4, Tl, seed=i, t_offset=t0_)
            if timing:
                ev[1].record()
            if world > 1:
                (am, pm), (aq, pq), _ = MD.allreduce_welford([r.attn_mean, r.prob_mean], [r.attn_m2, r.prob_m2], Tl,
                                                             total_count=T4)
            else:
                am, pm, aq, pq = r.attn_mean, r.prob_mean, r.attn_m2, r.prob_m2
            if timing:
                ev[2].record()
                ev[2].synchronize()           # per-step split (adds a host sync per step: reported separately below)
                split_ms[0] += ev[0].elapsed_time(ev[1])
                split_ms[1] += ev[1].elapsed_time(ev[2])
            merged[0] = (r, am, aq, pm, pq)

        def step4_nosync(i):
            r = mm.mc_head(w, H4, Tl, seed=i, t_offset=t0_)
            if world > 1:
                MD.allreduce_welford([r.attn_mean, r.prob_mean], [r.attn_m2, r.prob_m2], Tl, total_count=T4)

        steps4 = max(3, min(args.steps, 10))
        ms4, (p_ms, p_k) = timed(step4_nosync, steps4, 3, profile=True)
        timed(step4, steps4, 0)
        head_ms, merge_ms = split_ms[0] / steps4, split_ms[1] / steps4
        # ---- merged statistics vs ONE rank computing all T4 samples of the same Philox stream (every rank checks)
        seed_chk = 31337
        step4(seed_chk)
        r, am, aq, pm, pq = merged[0]
        full = mm.mc_head(w, H4, T4, seed=seed_chk)
        d_mean = float((am - full.attn_mean).abs().max())
        d_m2 = float((aq - full.attn_m2).abs().max() / full.attn_m2.abs().max())
        d_pm = float((pm - full.prob_mean).abs().max())
        d_pq = float((pq - full.prob_m2).abs().max() / max(float(full.prob_m2.abs().max()), 1e-30))
        y_equal = bool(torch.equal(r.Y, full.Y[:, t0_:t0_ + Tl]))
        ok = d_mean <= 1e-7 and d_m2 <= 1e-4 and d_pm <= 1e-6 and d_pq <= 1e-3 and y_equal
        flags = torch.tensor([1.0 if ok else 0.0, d_mean, d_m2, d_pm, d_pq], dtype=torch.float64, device=dev)
        if world > 1:
            mn = flags.clone()
            dist.all_reduce(mn, op=dist.ReduceOp.MIN)
            dist.all_reduce(flags, op=dist.ReduceOp.MAX)
            all_ok = bool(mn[0].item() == 1.0)
        else:
            all_ok = ok
        payload = (1 + 2 * (C * N4 + C)) * 8
        return {"scaling": "strong", "N": N4, "T": T4, "T_per_rank": Tl, "n_gpus": world, "steps": steps4,
                "ms_per_step": ms4 / steps4, "bags_per_s": steps4 / (ms4 / 1e3),
                "projection_ms_per_step_this_rank": p_ms / steps4,
                "head_ms_this_rank": head_ms, "merge_ms_this_rank": merge_ms,
                "collective": ("one NCCL all_reduce(sum) of %d bytes (fp64 additive Welford form) per step" % payload)
                if world > 1 else "none (single rank)",
                "merge_check": {"ok": all_ok, "ranks_checked": world,
                                "vs": "single-rank mc_head over all T samples of the same Philox stream, same seed",
                                "attn_mean_max_abs_diff": float(flags[1]), "attn_m2_max_diff_rel_to_max": float(flags[2]),
                                "prob_mean_max_abs_diff": float(flags[3]), "prob_m2_max_diff_rel_to_max": float(flags[4]),
                                "per_sample_logits_bit_equal": y_equal,
                                "bounds": "mean 1e-7 abs, M2 1e-4 of max (fp32 Welford grouping differs), prob mean 1e-6"}}

    def config5_leg():
        """BASELINE.json configs[4]: the infer.py path end to end on this GPU (infer.py:187-219 without DICOM loading and
        plotting): synthetic 2294x1914 image -> tiling / bag selection kernels -> torch ResNet-18 (batch-statistics
        BatchNorm, once per bag) -> fused MC head T=100 -> attention-map statistics per tile-boundary cell."""
        hh, ww, T5 = 2294, 1914, 100
        torch.manual_seed(0)
        model = mm.MultiHeadGatedAttentionMIL(pretrained=False, shared_attention=False)     # config.yml default
        model.apply(mm.deactivate_batchnorm)                  # infer.py:105-109,154
        model.to(dev).eval()
        model.extractor_mode = "channels_last"
        pt = mm.ImagePatcher(patch_size=224, overlap=0.75, bag_size=-1, empty_thresh=0.75)
        pt.get_tiles(hh, ww)
        g5 = torch.Generator(device=dev).manual_seed(0)
        yy = torch.arange(hh, device=dev).view(-1, 1).float()
        xx = torch.arange(ww, device=dev).view(1, -1).float()
        inside = ((yy - hh / 2) / (0.45 * hh)) ** 2 + (xx / (0.8 * ww)) ** 2 < 1.0
        img = ((torch.rand((1, hh, ww), generator=g5, device=dev) * 0.9 + 0.1) * inside).expand(3, hh, ww).contiguous()
        times = []
        for rep in range(4):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            ev[0].record()
            bag, idx, _ = pt.convert_img_to_bag(img)
            ev[1].record()
            with torch.no_grad():
                Hf = model.extract_features(bag.unsqueeze(0))
            ev[2].record()
            r5 = mm.mc_head(model._head_weights(dev), Hf, T5, seed=rep, return_attention=True)
            ev[3].record()
            st = pt.attention_map_stats(r5.A, idx, (hh, ww))
            mm_, sd_ = st.mean_map(), st.std_map()
            ev[4].record()
            torch.cuda.synchronize(dev)
            if rep > 0:
                times.append([ev[i].elapsed_time(ev[i + 1]) for i in range(4)])
        t5 = np.array(times).mean(0)
        return {"workload": f"{hh}x{ww} synthetic image, overlap 0.75: {len(idx)} of {len(pt.tiles)} tiles, ResNet-18 "
                            f"(torch, channels-last, batch-statistics BatchNorm), separate attention, T={T5}",
                "ms": {"tiling_and_bag": float(t5[0]), "resnet18_features_torch": float(t5[1]), "mc_head": float(t5[2]),
                       "attention_map_stats": float(t5[3]), "total": float(t5.sum())},
                "images_per_s": 1e3 / float(t5.sum()), "map_shape": list(mm_.shape),
                "reference_cpu_anchor": "BASELINE.md: tiling 0.67 s, attention-map reconstruction 5.9 s on the survey host"}

    config3 = config4 = config5 = None
    if extras_on and shared and not args.no_configs:
        config3 = config3_leg()
        config4 = config4_leg()
        if rank == 0 and world == 1:
            try:
                config5 = config5_leg()
            except Exception as e:  # noqa: BLE001   (torchvision missing on the box, ...)
                config5 = {"error": repr(e)}

    # ---- what a reference user with a GPU runs today: the reference's ATen op sequence on this B200 (informational)
    eager = None
    if extras_on and rank == 0 and world == 1:
        try:
            from oracle import torch_port as TP
            sd_dev = {k: v.to(dev) for k, v in make_state_dict(0, shared).items()}
            Hb = H[:1024]
            for _ in range(3):
                TP.mc_head_torch(sd_dev, Hb, 100, 0.1, 0.1)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for _ in range(20):
                TP.mc_head_torch(sd_dev, Hb, 100, 0.1, 0.1)
            torch.cuda.synchronize(dev)
            per = (time.perf_counter() - t0) / 20
            eager = {"bags_per_s": 1.0 / per, "ms_per_bag": per * 1e3,
                     "what": "oracle/torch_port.py (the reference's op sequence, fp32, torch eager, native CUDA dropout) on "
                             "this GPU, one bag of N=1024, T=100 per call; informational, not the reference arm"}
        except Exception as e:  # noqa: BLE001
            eager = {"error": repr(e)}

    # ---- CPU baseline on this host (rank 0, N=1 only): bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        Ncpu, Tcpu = (1024, 100) if args.workload == "config2" else (lens[0], min(T, 25))
        bps, times, head = cpu_head_bags_per_s(args.cpu_bags, Ncpu, Tcpu, shared, cores)
        cpu = {"value": bps * (Tcpu / (100 if args.workload == "config2" else T)), "unit": "bags/s", "cores": cores,
               "kind": head.kind,
               "sample": f"{args.cpu_bags} bag(s) of N={Ncpu}, T={Tcpu}: {head.describe()}, {cores} threads"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "bags/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f16xf16->f32",
            "data": "synthetic",
            "config": config_of(args.workload, len(all_lens) if strong else n_bags, R, merged_T[0] if strong else T,
                                C, shared, strong),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_f16_features": e2e_f16, "single_bag": single,
            "separate": separate, "config3": config3, "config4": config4, "config5": config5, "torch_cuda_eager": eager,
            "philox_rounds": args.philox_rounds, "philox7_mode": philox7,
            "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
            "clocks": clocks, "flops_per_step_per_gpu": flops_step, "hbm_peak_gbs": hbm_peak,
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
    ap.add_argument("--e2e-streams", type=int, default=3)
    ap.add_argument("--cpu-bags", type=int, default=8)
    ap.add_argument("--philox-rounds", type=int, default=10, choices=[7, 10])
    ap.add_argument("--no-extras", action="store_true", help="skip every extra leg (Philox-7, separate, single bag, configs 3/4)")
    ap.add_argument("--no-configs", action="store_true", help="skip the config 3 / config 4 legs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
