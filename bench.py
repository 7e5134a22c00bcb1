#!/usr/bin/env python
"""Headline benchmark: HR-Mpixels/s of the SRFlow-LP 4x inverse-sampling path (BASELINE.json config 2:
RRDB encoder, synthetic 160x160 LR tiles, batch 32 per GPU) on 1..8 B200.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference algorithm on the host CPU (oracle port, literal work)

One step = one pass of the whole LP path (encoder -> flow forward on bilinear(LR) -> latent normalisation ->
learned prior -> flow inverse) over the batch.  `value` is timed with CUDA events with the LR batch already resident in
HBM; `e2e` is the same call through the host-buffer C-ABI entry point (pinned host LR in, SR back to pinned host memory
every step).  Scaling: the default is STRONG -- the global batch of 32 tiles is split over the ranks with
`bfsr_b200.dist.shard_range` (32/N tiles per GPU, 4 at N = 8), no collective in the data path; the same run also times the
WEAK case (32 tiles per GPU) and reports it under `weak`.  `--scaling weak` swaps the two.  `--gather` adds the optional
final NCCL gather of the SR tiles to rank 0 inside the timed region.
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

import torch  # noqa: E402

SCALE, LR_SIZE, BATCH = 4, 160, 32
# deduplicated algorithmic work of the path, SURVEY.md §8(d): 57.094 M MAC per LR pixel
ALG_FLOP_PER_LR_PX = 2 * 57.094e6


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            tf = d.get("bf16_tflops_sustained") or d["bf16_tflops"]
            return {"hbm_gbs": float(d["hbm_gbs"]), "tflops": float(tf), "src": "measured"}
        except Exception as e:   # unreadable / unexpected layout: say so and use the documented fallback
            sys.stderr.write(f"bench: MEASURED_PEAKS.json not usable ({e!r}); using the fallback peaks\n")
    # B200_PROFILING.md fallback: 6.65 TB/s copy, 1.59 PFLOP/s bf16 burst, ~1.4 PFLOP/s sustained under the 1 kW cap; the step is a
    # ~0.3 s stream of tensor-core kernels, so the sustained figure is the denominator
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "src": "fallback (sustained bf16 of B200_PROFILING.md)"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons for one GPU during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def _top_launch_traffic():
    """dram__bytes_read+write of the top single launch from the committed `ncu --set full` capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "r1_ftconv_full_summary.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


def workload(topo_kw=None):
    from tools import synth
    t = synth.SRFlowTopo(**(topo_kw or {}))
    sd = synth.synth_srflow_state_dict(t, seed=0)
    usd = synth.synth_unet_state_dict(synth.unet_srflow_param_shapes(), seed=1)
    return t, sd, usd


def cpu_reference_step(t, sd, usd, lr_tile):
    """The reference's algorithm for the path (literal work: encoder in both passes, dead heads) on the host CPU."""
    from oracle import srflow_oracle as O
    t0 = time.perf_counter()
    sr = O.lp_sr(sd, usd, t, lr_tile, literal=True)
    dt = time.perf_counter() - t0
    return dt, sr


def run_reference(args, rank, world):
    if rank != 0:
        return
    from tools import synth
    torch.set_num_threads(os.cpu_count() or 1)
    t, sd, usd = workload()
    lr = synth.img(1, LR_SIZE, LR_SIZE, 1234 + 2)
    for _ in range(args.warmup):
        cpu_reference_step(t, sd, usd, lr[:, :, :40, :40].contiguous())   # warm the host libraries on a small tile
    times = [cpu_reference_step(t, sd, usd, lr)[0] for _ in range(args.steps)]
    ms = 1e3 * sum(times) / len(times)
    hr_px = (SCALE * LR_SIZE) ** 2
    val = hr_px / (ms * 1e-3) / 1e6
    sample = f"1 tile of {LR_SIZE}x{LR_SIZE} LR per step (of the {BATCH}-tile batch), literal reference work, fp32"
    print(json.dumps({
        "impl": "reference", "metric": "HR Mpixels/sec, SRFlow-LP 4x LP inference (160x160 LR tiles)", "value": val,
        "unit": "HR-Mpix/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"SRFlow-LP 4x RRDB(nb=23,K=16,L=3) {LR_SIZE}x{LR_SIZE} LR tiles, batch {BATCH} per GPU, synthetic weights",
                   "sample": sample},
        "cpu_baseline": {"value": val, "unit": "HR-Mpix/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": val, "unit": "HR-Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="bfsr", choices=["bfsr", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--lr-size", type=int, default=LR_SIZE)
    ap.add_argument("--tile-chunk", type=int, default=0)
    ap.add_argument("--precision", type=int, default=int(os.environ.get("BFSR_PRECISION", "0")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--gather", action="store_true", help="time the optional final gather of the SR tiles to rank 0 (NCCL) too")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from bfsr_b200 import _lib, models
    from tools import synth

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout to the single JSON line: NCCL prints its version banner (NCCL_DEBUG=VERSION in this pool) to stdout while
        # the communicator is created, so fd 1 points at stderr until the first collective has run
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    warm = max(args.warmup, 3)
    B, S = args.batch, args.lr_size

    t, sd, usd = workload()
    net = models.define_Flow(t.opt(), device=dev, tile_chunk=args.tile_chunk, precision=args.precision)
    net.load_state_dict(sd, strict=True)
    prior = models.make({"name": "unet", "args": {"depth": 3, "dim": 64, "bilinear": True}, "sd": usd}, load_sd=True)
    from bfsr_b200.dist import gather_tiles, shard_range
    # the global batch (rank 0's seed, so tile 0 is the tile of the reference fixture) and this rank's share of it
    lr_global = synth.img(B, S, S, 1234 + 2)
    lo, hi = shard_range(B, world, rank) if args.scaling == "strong" else (0, B)
    if args.scaling == "weak" and rank:
        lr_global = synth.img(B, S, S, 1234 + 2 + rank)
    B_local = hi - lo
    lr_host = lr_global[lo:hi].contiguous().pin_memory()
    lr = lr_host.to(dev)
    L = _lib.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        tt = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def timed(x, steps, gather=False):
        """max over ranks of the CUDA-event time of `steps` passes over x, bracketed by barrier + synchronize"""
        out = None
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = net.lp_sr(x, prior)
            if gather and world > 1:
                gather_tiles(out, B if args.scaling == "strong" else B * world, dst=0)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / steps, out

    # ---- device-resident timing
    for _ in range(warm):
        sr = net.lp_sr(lr, prior)
        if args.gather and world > 1:
            gather_tiles(sr, B if args.scaling == "strong" else B * world, dst=0)
    barrier()
    L.bfsr_launch_count(1)
    with ClockSampler(local) as cs:
        ms, sr = timed(lr, args.steps, gather=args.gather)
    launches = int(L.bfsr_launch_count(0))
    clocks = cs.summary()
    tiles_job = B if args.scaling == "strong" else B * world
    value = tiles_job * (SCALE * S) ** 2 / (ms * 1e-3) / 1e6
    assert torch.isfinite(sr).all()
    # the other scaling mode in the same run (N > 1 only; at N = 1 they coincide)
    other = None
    if world > 1:
        if args.scaling == "strong":
            x2 = synth.img(B, S, S, 1234 + 2 + rank).to(dev)        # weak: 32 tiles per GPU
            tiles2 = B * world
        else:
            l2, h2 = shard_range(B, world, rank)
            x2 = synth.img(B, S, S, 1234 + 2)[l2:h2].contiguous().to(dev)
            tiles2 = B
        for _ in range(2):
            net.lp_sr(x2, prior)
        ms2, _ = timed(x2, args.steps)
        other = {"scaling": "weak" if args.scaling == "strong" else "strong", "value": tiles2 * (SCALE * S) ** 2 / (ms2 * 1e-3) / 1e6,
                 "unit": "HR-Mpix/s", "ms_per_step": ms2, "global_batch": tiles2, "tiles_per_gpu": tiles2 // world}
        del x2

    # ---- end to end through the host-buffer C-ABI entry point
    out_host = torch.empty((B_local, 3, SCALE * S, SCALE * S), dtype=torch.float32).pin_memory()
    net.lp_sr_host(lr_host, prior, out=out_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        net.lp_sr_host(lr_host, prior, out=out_host)     # returns after SR is in host memory
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    e2e = {"value": tiles_job * (SCALE * S) ** 2 / (e2e_ms * 1e-3) / 1e6, "unit": "HR-Mpix/s",
           "h2d_bytes_per_step": lr_host.numel() * 4 * world, "d2h_bytes_per_step": out_host.numel() * 4 * world,
           "ms_per_step": e2e_ms}

    # ---- per-kernel-class device time of one more step (CUDA events around every launch, on the launching stream)
    import ctypes as C
    peaks = load_peaks()
    roofline = None
    L.bfsr_prof_enable(1)
    net.lp_sr(lr, prior)
    cls = {}
    for kind, name in ((0, "conv_fp32"), (1, "conv_tcgen05"), (2, "flowstep"), (3, "other")):
        tms, work, cnt = C.c_double(), C.c_double(), C.c_int64()
        L.bfsr_prof_summary(kind, C.byref(tms), C.byref(work), C.byref(cnt))
        cls[name] = {"ms": tms.value, "work": work.value, "launches": cnt.value}
    L.bfsr_prof_enable(0)
    conv_name = max(("conv_fp32", "conv_tcgen05"), key=lambda k: cls[k]["ms"])
    cv = cls[conv_name]
    if cv["ms"] > 0:
        ach = cv["work"] / (cv["ms"] * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": conv_name, "achieved": ach, "peak": peaks["tflops"], "unit": "TFLOP/s",
                    "frac": ach / peaks["tflops"], "traffic": None, "peak_source": peaks["src"],
                    "launches": cv["launches"], "kernel_ms_per_step": cv["ms"],
                    "share_of_step": cv["ms"] / ms if ms else None,
                    "peak_3pass": peaks["tflops"] / 3.0, "frac_of_3pass_peak": ach / (peaks["tflops"] / 3.0),
                    "note": "all tcgen05 conv launches of one step: algorithmic conv FLOPs (2*px*Cin*k*k*Cout of the conv the "
                            "reference executes, unpadded) / summed device time (CUDA events around every launch); peak = dense "
                            "bf16 sustained; the parity mode spends 3 bf16 MMAs per product (split-bf16 x3), so peak_3pass is "
                            "the ceiling of this arithmetic"}
        top = _top_launch_traffic()
        if top:
            roofline["traffic"] = top.get("dram_bytes_per_launch")
            roofline["traffic_kernel"] = top.get("kernel")
            roofline["traffic_algorithmic_bytes"] = top.get("algorithmic_bytes_per_launch")
    fs = cls["flowstep"]
    flow_roof = None
    if fs["ms"] > 0:
        ach = fs["work"] / (fs["ms"] * 1e-3) / 1e9
        flow_roof = {"bound": "hbm", "kernel": "flowstep", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": ach / peaks["hbm_gbs"], "launches": fs["launches"], "kernel_ms_per_step": fs["ms"]}

    # ---- CPU baseline (rank 0, N=1): the oracle port on a bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        cpu_reference_step(t, sd, usd, lr_host[:1, :, :40, :40].contiguous())
        dt, ref = cpu_reference_step(t, sd, usd, lr_host[:1].contiguous())
        got = sr[:1].cpu().double()
        rel = float((got - ref.double()).norm() / ref.double().norm())
        cpu = {"value": (SCALE * S) ** 2 / dt / 1e6, "unit": "HR-Mpix/s", "cores": torch.get_num_threads(),
               "kind": "port", "sample": f"1 of {tiles_job} tiles ({S}x{S} LR), literal reference work, {dt:.1f} s",
               "parity_rel_l2_vs_gpu": rel}

    # ---- parity of this very run against the UNMODIFIED reference: tile 0 of rank 0's batch is the tile recorded in
    # tests/golden/srflow_full160.npz by oracle/make_golden_r2.py (stride-4 grid of the reference's SR output, pre-clamp)
    ref_parity = None
    if rank == 0 and B >= 1 and S == LR_SIZE:
        try:
            import numpy as np
            g = np.load(os.path.join(ROOT, "tests", "golden", "srflow_full160.npz"))
            ref_s4 = torch.from_numpy(g["sr_s4"]).double()
            got_s4 = sr[:1, :, ::4, ::4].cpu().double()
            rel = float((got_s4 - ref_s4).norm() / ref_s4.norm())
            a, b = got_s4.clamp(0, 1), ref_s4.clamp(0, 1)
            mse = float(((a - b) ** 2).mean())
            ref_parity = {"rel_l2_preclamp": rel, "psnr_db_clamped": (float("inf") if mse == 0 else -10.0 * __import__("math").log10(mse)),
                          "fixture": "tests/golden/srflow_full160.npz (reference run of tile 0, stride-4 grid)"}
        except Exception as e:     # fixture missing: say so instead of inventing a number
            ref_parity = {"unavailable": repr(e)}

    if rank == 0:
        print(json.dumps({
            "metric": "HR Mpixels/sec, SRFlow-LP 4x LP inference (160x160 LR tiles)", "value": value, "unit": "HR-Mpix/s",
            "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32" if args.precision == 0 else "bf16", "data": "synthetic",
            "config": {"workload": f"SRFlow-LP 4x RRDB(nb=23,K=16,L=3) {S}x{S} LR tiles, global batch {tiles_job}, synthetic weights",
                       "global_batch": tiles_job, "tiles_per_gpu": B_local, "gather_in_timed_region": bool(args.gather and world > 1), "precision_mode": {0: "fp32-accurate (split-bf16 x3 on tcgen05, fp32 accumulate)", 1: "bf16-fast", 2: "fp32 CUDA cores"}[args.precision],
                       "l2": "per-step working set (tens of GB of activations) >> 126 MB L2, no flush needed",
                       "parallelism": f"dp{world} (independent tiles, no data-path collective)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "roofline_flowstep": flow_roof,
            "kernel_classes_ms_per_step": {k: round(v["ms"], 3) for k, v in cls.items()},
            "alg_tflop_per_step": ALG_FLOP_PER_LR_PX * tiles_job * S * S / 1e12,
            "path_tensor_roofline_frac": (ALG_FLOP_PER_LR_PX * tiles_job * S * S / world / (ms * 1e-3) / 1e12) / peaks["tflops"],
            "workspace_gb": L.bfsr_srflow_workspace_bytes(net.handle()) / 1e9,
            "cpu_baseline": cpu, "parity_vs_reference": ref_parity, "weak" if args.scaling == "strong" else "strong": other,
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
