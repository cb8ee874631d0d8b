#!/usr/bin/env python
"""bench.py -- output pixels/s of UMPAModelDF.match() (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2]

One "step" = one match() over one synthetic speckle stack of the named configuration.
N > 1 is launched with torchrun (one rank per GPU); the output rows of the SAME stack are
split into N row bands (+ halo), no collective on the data path (strong scaling).

Printed JSON (one line, rank 0):
  value        device-resident throughput: centred FP32 stacks already in HBM, result maps left
               in HBM; CUDA events around K steps, max over ranks.
  e2e          same metric through the public drop-in API with HOST buffers: per step the float64
               stacks go host(pinned)->device, are converted, matched, and the result maps come
               back to pinned host memory.
  roofline     dominant kernel (cross-correlation shift table): ALGORITHMIC flops
               2*S^2*Na*K^2 per output pixel / measured kernel time, against the
               FP32-FMA peak measured on this GPU in this run by an FFMA probe.  The kernel
               executes ~K^2/ (1+2K/Na) times fewer FMAs than that (frame-sum first, filter once),
               so `frac` can exceed 1; `executed_*` report what the SMs really did.
  cpu_baseline the reference's own OpenMP CPU path (oracle/_ref, built from /root/reference) or
               the C port, timed on this box's host cores on a bounded ROI of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIGS = {
    "cfg1": dict(kind="NoDF", Na=10, H=256, W=256, Nw=2, ms=4, desc="UMPAModelNoDF 10x256^2 Nw=2 max_shift=4"),
    "cfg2": dict(kind="DF", Na=25, H=2048, W=2048, Nw=2, ms=5, desc="UMPAModelDF 25x2048^2 Nw=2 max_shift=5"),
    "cfg3": dict(kind="DFKernel", Na=25, H=2048, W=2048, Nw=3, ms=5, desc="UMPAModelDFKernel 25x2048^2 Nw=3 max_shift=5"),
    "cfg4": dict(kind="DF", Na=40, H=4096, W=4096, Nw=3, ms=8, desc="UMPAModelDF 40x4096^2 Nw=3 max_shift=8"),
    "cfg5": dict(kind="NoDF", Na=4, H=2048, W=2048, Nw=6, ms=4, desc="UMPAModelNoDF 4x2048^2 Nw=6 max_shift=4"),
}
SAFE_CROP = {"NoDF": 0, "DF": 0, "DFKernel": 8}
# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed
# `ncu --set full` capture of the same command (profiles/r01f_final_kernels_ncu_summary.txt): the cross-table
# kernel reads the two FP32 stacks once (0.88 GB with the tile halos) and writes the 81-plane table (1.31 GB).
NCU_TRAFFIC = {("cfg2", 1): 880.3e6 + 1.3138e9}
NCU_TRAFFIC_SOURCE = "profiles/r01f_final_kernels_ncu_summary.txt (ncu --set full, shift_table_kernel<9,3,2>)"


def algorithmic_flops_per_px(cfg):
    """F_alg = 2*S^2*Na*K^2 (+ blur for DFKernel)."""
    S, K, Na = 2 * cfg["ms"] - 1, 2 * cfg["Nw"] + 1, cfg["Na"]
    f = 2. * S * S * Na * K * K
    if cfg["kind"] == "DFKernel":
        f = 2. * Na * ((K + S - 1) ** 2 * 17 ** 2 + 2 * S * S * K * K)
    return f


def executed_fma_per_px_cross(cfg):
    """FMAs the cross-table kernel executes per output pixel: Na per (extended-tile pixel, shift)
    plus the separable filter (row pass on the extended rows, column pass).  Tile geometry of
    table_path.cu: extended tile 16 x 32, output tile (16-2Nw) x (32-2Nw)."""
    S, K, Na, Nw = 2 * cfg["ms"] - 1, 2 * cfg["Nw"] + 1, cfg["Na"], cfg["Nw"]
    # extended tile height as plan_tiles (table_path.cu) picks it
    def _cost(eh):
        sh = 3 if S <= 9 else (2 if S <= 17 else 1)
        g = min(384 // (eh * 8), -(-S // sh))
        return eh / float(eh - 2 * Nw) * (1. + .15 * (-(-S // (g * sh)) - 1))
    eh = min((e for e in (16, 24, 32, 48) if e - 2 * Nw >= 2), key=lambda e: (_cost(e), e))
    eh, ew = int(os.environ.get("UMPA_TAB_EH", eh)), 32
    th, tw = eh - 2 * Nw, ew - 2 * Nw
    per_tile = S * S * (Na * eh * ew + K * eh * tw + K * th * tw)
    return per_tile / float(th * tw)


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (NVML)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t is not None:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def make_stacks(cfg, device, seed=2):
    from umpa_b200 import synth
    d = synth.speckle_stack(cfg["Na"], cfg["H"], cfg["W"], seed=seed, max_shift=cfg["ms"],
                            dark_field=cfg["kind"] != "NoDF", device=device, as_numpy=False)
    return d["sam"], d["ref"]


# ------------------------------------------------------------------------------ reference arm

def cpu_model(cfg, sam_np, ref_np):
    """(model factory result, kind): the compiled reference when available, else the C port."""
    from oracle import ref as oref
    R = oref.load(build_if_missing=os.path.exists("/root/reference"))
    sam_l, ref_l = [s for s in sam_np], [r for r in ref_np]
    if R is not None:
        cls = {"NoDF": R.UMPAModelNoDF, "DF": R.UMPAModelDF, "DFKernel": R.UMPAModelDFKernel}[cfg["kind"]]
        return cls(sam_l, ref_l, window_size=cfg["Nw"], max_shift=cfg["ms"]), "reference"
    from oracle import port
    return port.OracleModel(cfg["kind"], sam_l, ref_l, window_size=cfg["Nw"], max_shift=cfg["ms"]), "port"


def cpu_match(model, kind, roi, cores, cfg):
    kw = {}
    if cfg["kind"] == "DFKernel":
        from umpa_b200 import synth
        n0 = 1 + (roi[0][1] - roi[0][0] - 1) // roi[0][2]
        n1 = 1 + (roi[1][1] - roi[1][0] - 1) // roi[1][2]
        kw["abc"] = synth.blur_abc(n0, n1)
    t = time.perf_counter()
    if kind == "reference":
        res = model.match(ROI=roi, num_threads=cores, quiet=True, **kw)
    else:
        res = model.match(ROI=roi, num_threads=cores, debug=False, **kw)
    dt = time.perf_counter() - t
    return res["err"].size, dt


def centred_roi(cfg, n_px_target):
    pad = cfg["ms"] + cfg["Nw"] + SAFE_CROP[cfg["kind"]]
    N0, N1 = cfg["H"] - 2 * pad, cfg["W"] - 2 * pad
    rows = int(max(8, min(N0, round(n_px_target / float(N1)))))
    r0 = (N0 - rows) // 2
    return ((r0, r0 + rows, 1), (0, N1, 1)), rows * N1


def run_cpu_sample(cfg, sam_np, ref_np, target_s, steps=1, warmup=0):
    """Times the CPU path on a centred full-width row block sized for ~target_s per step."""
    cores = os.cpu_count() or 1
    model, kind = cpu_model(cfg, sam_np, ref_np)
    roi, npx = centred_roi(cfg, 20000)
    _, dt = cpu_match(model, kind, roi, cores, cfg)                 # calibration (also warms caches)
    rate = npx / max(dt, 1e-6)
    roi, npx = centred_roi(cfg, rate * target_s)
    for _ in range(warmup):
        cpu_match(model, kind, roi, cores, cfg)
    times = []
    for _ in range(steps):
        n, dt = cpu_match(model, kind, roi, cores, cfg)
        times.append(dt)
    t = float(np.mean(times))
    sample = "centred full-width block of %d rows (%d px) of %s, %d threads" % (
        roi[0][1] - roi[0][0], npx, cfg["desc"], cores)
    return dict(value=npx / t, unit="output pixels/s", cores=cores, kind=kind, sample=sample), t, npx


def reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    sam, ref = make_stacks(cfg, dev)
    sam_np, ref_np = sam.cpu().numpy(), ref.cpu().numpy()
    del sam, ref
    # each step ~ (120 s budget) / (steps + warmup + calibration)
    per_step = max(1.0, min(15.0, 120.0 / (args.steps + args.warmup + 1)))
    base, t, npx = run_cpu_sample(cfg, sam_np, ref_np, per_step, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": "output pixels/s", "value": base["value"], "unit": "output pixels/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": cfg["desc"], "sample": base["sample"]},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "output pixels/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ our arm

def ours(args, cfg):
    import torch
    import torch.distributed as dist
    import umpa_b200
    from umpa_b200 import _capi
    from umpa_b200.sharding import ShardedMatcher
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    cls = {"NoDF": umpa_b200.UMPAModelNoDF, "DF": umpa_b200.UMPAModelDF,
           "DFKernel": umpa_b200.UMPAModelDFKernel}[cfg["kind"]]
    sam, ref = make_stacks(cfg, dev)
    pad = cfg["ms"] + cfg["Nw"] + SAFE_CROP[cfg["kind"]]
    N0, N1 = cfg["H"] - 2 * pad, cfg["W"] - 2 * pad
    total_px = N0 * N1
    kw = {}
    if cfg["kind"] == "DFKernel":
        from umpa_b200 import synth
        kw["abc"] = synth.blur_abc(N0, N1, as_numpy=False).to(dev)

    # ---- device-resident metric -------------------------------------------------
    sm = ShardedMatcher(cls, list(sam), list(ref), rank, world, window_size=cfg["Nw"], max_shift=cfg["ms"])
    band_rows = sm.band[1] - sm.band[0]
    lo, hi = sm.band[0], sm.band[1] + 2 * pad
    launches = 0
    for _ in range(args.warmup):
        out = sm.match_device(**kw)
    info = sm.model.last_match_info if sm.model is not None else {"path": "none", "kernel_launches": 0}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with ClockSampler(local) as clk:
        e0.record()
        for _ in range(args.steps):
            out = sm.match_device(**kw)
            launches += info["kernel_launches"]
        e1.record()
        torch.cuda.synchronize()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    value = total_px / (ms * 1e-3)
    launches = int(sum_over_ranks(launches))
    err_ok = float(sum_over_ranks(float((out["err"] == 1).sum().item()) if out else 0.)) / total_px

    # ---- stage times of the dominant kernel (events inside the library, same stream) -------
    stage_ms = None
    if sm.model is not None and info["path"] == "table":
        _capi.check(_capi.lib().umpa_set_profiling(sm.model._h, 1))
        acc = np.zeros(4)
        reps = max(3, min(args.steps, 10))
        for _ in range(reps):
            sm.match_device(**kw)
            buf = (C.c_float * 4)()
            n = _capi.lib().umpa_last_stage_ms(sm.model._h, buf, 4)
            acc += np.array(buf[:4]) if n == 4 else 0
        _capi.check(_capi.lib().umpa_set_profiling(sm.model._h, 0))
        stage_ms = (acc / reps).tolist()
    peak = C.c_double(0.)
    sms = C.c_int(0)
    _capi.check(_capi.lib().umpa_fma_peak(C.byref(peak), C.byref(sms)))

    # ---- end to end through the drop-in API with host buffers --------------------------------
    e2e = None
    if not args.no_e2e:
        hs = torch.empty((cfg["Na"], hi - lo, cfg["W"]), dtype=torch.float64, pin_memory=True)
        hr = torch.empty_like(hs, pin_memory=True)
        hs.copy_(sam[:, lo:hi]); hr.copy_(ref[:, lo:hi])
        torch.cuda.synchronize()
        sam_np, ref_np = hs.numpy(), hr.numpy()
        kw_h = {}
        if "abc" in kw:
            kw_h["abc"] = kw["abc"][sm.band[0]:sm.band[1]].cpu().numpy()
        del sm
        steps_e = max(2, min(args.steps, 10))
        d2h = 0

        stream_info = {}

        def one():
            m = cls(list(sam_np), list(ref_np), window_size=cfg["Nw"], max_shift=cfg["ms"])
            r = m.match(quiet=True, debug=False, **kw_h)
            stream_info.update(m.last_stream_info)
            return sum(v.nbytes for v in r.values())
        for _ in range(max(3, args.warmup)):     # (the first calls pin staging / result buffers and measure the host rates)
            d2h = one()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps_e):
            d2h = one()
        torch.cuda.synchronize()
        t_e = (time.perf_counter() - t0) / steps_e
        barrier()
        t_e = max_over_ranks(t_e)
        e2e = {"value": total_px / t_e, "unit": "output pixels/s", "ms_per_step": 1e3 * t_e,
               "h2d_bytes_per_step": int(sum_over_ranks(float(sam_np.nbytes + ref_np.nbytes))),
               "d2h_bytes_per_step": int(sum_over_ranks(float(d2h))), "steps": steps_e,
               "api": "%s(sam, ref, window_size, max_shift).match(debug=False) on pinned host float64 arrays"
                      % cls.__name__,
               "pipeline": dict(stream_info, note="upload / kernels / download overlapped in row bands; host_threads "
                                "convert the lower host_rows_per_frame rows of every frame to centred FP32 in pinned "
                                "staging while the DMA engine uploads the upper rows as FP64 (UMPA_HOST_THREADS=0 "
                                "disables the host part)")}
        # the same call on float32 host frames (detector data; umpa_set_frames_f32): half the upload, no host
        # conversion.  Reported beside the headline, not instead of it: the reference's API is float64.
        hs32 = torch.empty(hs.shape, dtype=torch.float32, pin_memory=True)
        hr32 = torch.empty(hs.shape, dtype=torch.float32, pin_memory=True)
        hs32.copy_(hs); hr32.copy_(hr)
        s32, r32 = hs32.numpy(), hr32.numpy()

        def one32():
            m = cls(list(s32), list(r32), window_size=cfg["Nw"], max_shift=cfg["ms"])
            m.match(quiet=True, debug=False, **kw_h)
        for _ in range(max(3, args.warmup)):
            one32()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps_e):
            one32()
        torch.cuda.synchronize()
        t_32 = (time.perf_counter() - t0) / steps_e
        barrier()
        t_32 = max_over_ranks(t_32)
        e2e["float32_frames"] = {"value": total_px / t_32, "ms_per_step": 1e3 * t_32,
                                 "h2d_bytes_per_step": int(sum_over_ranks(float(s32.nbytes + r32.nbytes))),
                                 "note": "same call, frames given as float32 numpy arrays (widened and centred on the GPU)"}

    # ---- CPU baseline (rank 0, N=1 only) --------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu, _, _ = run_cpu_sample(cfg, sam.cpu().numpy(), ref.cpu().numpy(), args.cpu_seconds)
        except Exception as e:      # the baseline must never take the GPU number down with it
            cpu = {"value": None, "unit": "output pixels/s", "cores": os.cpu_count(), "kind": "unavailable",
                   "sample": "failed: %r" % (e,)}

    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.)
    roof = None
    if stage_ms is not None:
        px_rank = band_rows * N1                         # rank 0's launch processes its band
        t_cross = stage_ms[1] * 1e-3
        f_alg = algorithmic_flops_per_px(cfg)
        dfk = cfg["kind"] == "DFKernel"
        if dfk:     # the blur-table kernel executes the direct form: patch blur + t5 partials + q3 (kernel_path.cu)
            S_, K_ = 2 * cfg["ms"] - 1, 2 * cfg["Nw"] + 1
            f_exe = 2. * cfg["Na"] * ((K_ + S_ - 1) ** 2 * 17 ** 2 + (K_ + S_ - 1) * K_ * K_ * S_ + (K_ + S_ - 1) ** 2)
            kname = "ktable_kernel<Nw=%d,S=%d> (per-pixel blur tables)" % (cfg["Nw"], S_)
            note = ("achieved = SURVEY 8d direct-form flop/px x px / kernel time; the kernel executes that form "
                    "(blur of the (K+S-1)^2 patch per frame; executed/algorithmic <= %.2f: kernel taps below 1e-10 of "
                    "the pixel's largest tap are zero and the blur loops stop at the last non-zero row / column)")
        else:
            f_exe = 2. * executed_fma_per_px_cross(cfg)
            kname = "shift_table_kernel<S=%d,Nw=%d> (cross table)" % (2 * cfg["ms"] - 1, cfg["Nw"])
            note = ("achieved = 2*S^2*Na*K^2 flop/px (SURVEY 8d, direct form) x px / kernel time; the kernel "
                    "sums over frames first and filters once, so it executes %.0fx fewer FMAs")
        ach = f_alg * px_rank / t_cross / 1e12
        exe = f_exe * px_rank / t_cross / 1e12
        roof = {"bound": "fp32_fma", "kernel": kname,
                "achieved": ach, "peak": peak.value, "unit": "TFLOP/s", "frac": ach / peak.value,
                "peak_source": "FFMA probe measured in this run on this GPU (umpa_fma_peak); nominal %.1f"
                               % (sms.value * 128 * 2 * (peaks.get("sm_max_mhz", 1965.) * 1e6) / 1e12),
                "traffic": NCU_TRAFFIC.get((args.config, world)), "traffic_source": NCU_TRAFFIC_SOURCE
                if (args.config, world) in NCU_TRAFFIC else None, "kernel_ms": stage_ms[1],
                "executed_tflops": exe, "executed_frac": exe / peak.value,
                "note": note % ((f_exe / f_alg) if dfk else (f_alg / f_exe)),
                "stage_ms": {"moments": stage_ms[0], "blur_table" if dfk else "cross_table": stage_ms[1],
                             "mean_table": stage_ms[2], "walk": stage_ms[3]}}
    alg_bytes = 2. * cfg["Na"] * cfg["H"] * cfg["W"] * 4 + 6 * 4. * total_px
    line = {"metric": "output pixels/s", "value": value, "unit": "output pixels/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["desc"], "output_px": total_px, "sharding": "row bands x%d + %d-row halo" % (world, pad),
                       "path": info["path"], "l2": "inputs (%.0f MB FP32 stacks) larger than the 126 MB L2, no flush" % (alg_bytes / 1e6),
                       "inputs_resident": "mean-centred FP32 stacks in HBM; result maps (f,T,dx,dy,df f64; err,Ncalls i32) left in HBM",
                       "err_ok_fraction": err_ok},
            "clocks": clk.summary(), "e2e": e2e, "gpu_launches": launches, "roofline": roof,
            "roofline_hbm_step": {"bound": "hbm", "achieved": alg_bytes / (ms * 1e-3) / 1e9 / world, "peak": hbm_peak,
                                  "unit": "GB/s", "frac": alg_bytes / (ms * 1e-3) / 1e9 / world / hbm_peak,
                                  "note": "algorithmic bytes of the whole step (stacks once + 6 maps) per GPU-second"},
            "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        reference_arm(args, cfg)
    else:
        ours(args, cfg)


if __name__ == "__main__":
    main()
