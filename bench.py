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
    p = os.path.join(ROOT, "profiles", "r2_ftconv_full_summary.json")
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
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"SRFlow-LP 4x RRDB(nb=23,K=16,L=3) {LR_SIZE}x{LR_SIZE} LR tiles, batch {BATCH} per GPU, synthetic weights",
                   "sample": sample},
        "cpu_baseline": {"value": val, "unit": "HR-Mpix/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": val, "unit": "HR-Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# =========================================================================================== BASELINE configs 3 / 4 / 5
def _barrier(dist, world, dev):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


def _max_over_ranks(dist, world, dev, x):
    if world == 1:
        return x
    tt = torch.tensor([x], device=dev, dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return float(tt.item())


def _linf_models(enc, dev, precision):
    """Shipped checkpoint when it has been exported next to the fixtures (tests/golden/_linf_ckpt, real weights), else synthetic
    weights of the same architecture (timing is data-independent)."""
    from bfsr_b200 import models
    from tools import synth
    path = os.path.join(ROOT, "tests", "golden", "_linf_ckpt", enc + ".pt")
    if os.path.exists(path):
        ck = torch.load(path, map_location="cpu")
        sd, psd, src = ck["model"]["sd"], ck["prior_model"]["sd"], "shipped checkpoint"
    else:
        sd = synth.synth_linf_state_dict(synth.linf_param_shapes(enc), seed=5)
        psd = synth.synth_unet_state_dict(synth.unet_linf_param_shapes(), seed=6)
        src = "synthetic weights"
    spec = {"name": "linf-patch", "args": {"encoder_spec": {"name": enc, "args": {"no_upsampling": True}},
                                           "imnet_spec": {"name": "flow", "args": {"name": "flow"}}, "flow_layers": 10,
                                           "num_layer": 3, "hidden_dim": 256, "patch_size": 3}, "sd": sd}
    model = models.make(spec, args={"precision": precision}, load_sd=True).cuda(dev)
    prior = models.make({"name": "unet", "args": {"in_chans": 27, "depth": 3, "dim": 64, "cell_input": False, "bilinear": True},
                         "sd": psd}, load_sd=True).cuda(dev)
    return model, prior, sd, psd, src


def run_linf(args, rank, world, local, cfg):
    """Config 3: LINF-LP EDSR-baseline 4x, 64 patches of 48x48 LR on 1 GPU (paired-wrapper inputs, q = 65).
    Config 5: LINF-LP RRDB, 256 patches of 48x48 LR with scales {2,3,4,6,8} cycling, every scale bucket sharded over the ranks."""
    import ctypes as C
    import torch.distributed as dist
    from bfsr_b200 import _lib, models
    from bfsr_b200.dist import bucket_by_scale, shard_range
    from tools import synth
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        saved_fd = os.dup(1); os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev); dist.barrier(); torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush(); os.dup2(saved_fd, 1); os.close(saved_fd)
    enc = "edsr-baseline" if cfg == 3 else "rrdb"
    B = args.batch if args.batch != BATCH else (64 if cfg == 3 else 256)
    model, prior, sd, psd, wsrc = _linf_models(enc, dev, args.precision)
    L = _lib.lib()
    lr01 = synth.img(B, 48, 48, 1234 + cfg)
    scales = [4] * B if cfg == 3 else [[2, 3, 4, 6, 8][i % 5] for i in range(B)]
    mine = bucket_by_scale(scales, world, rank) if cfg == 5 else {4: list(range(*shard_range(B, world, rank)))}
    # inputs are built once on the device (the wrappers' work, b7) and stay resident: the timed step is the model path
    buckets = []
    for s_, idx in mine.items():
        if idx:
            buckets.append((s_, idx, models.build_inputs(lr01[idx].to(dev), s_, 3, cfg == 3)))
    hr_px_job = sum((48 * s_) ** 2 for s_ in scales)

    def step():
        out = None
        for s_, idx, (inp, coord, cell, gt, hw) in buckets:
            out = model.lp_sr(inp, coord, cell, gt, prior, hw)
        return out

    warm = max(args.warmup, 3)
    for _ in range(warm + 2):                       # two extra passes: the third identical call onwards replays the CUDA graph
        out = step()
    _barrier(dist, world, dev)
    L.bfsr_launch_count(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2: written between timed steps
    times = []
    with ClockSampler(local) as cs:
        for _ in range(args.steps):
            flush.fill_(1)
            _barrier(dist, world, dev)
            e0.record(); out = step(); e1.record()
            _barrier(dist, world, dev)
            times.append(_max_over_ranks(dist, world, dev, e0.elapsed_time(e1)))
    launches = int(L.bfsr_launch_count(0)) // max(args.steps, 1)
    ms = sum(times) / len(times)
    value = hr_px_job / (ms * 1e-3) / 1e6
    assert out is None or torch.isfinite(out).all()
    # ---- e2e: host buffers in, SR back on the host, every step
    host = [(s_, idx, tuple(t.cpu().contiguous().pin_memory() for t in ins[:4]), ins[4]) for s_, idx, ins in buckets]
    outs_h = [torch.empty((len(idx), 3, hw[0], hw[1]), dtype=torch.float32).pin_memory() for s_, idx, t4, hw in host]

    def step_host():
        for (s_, idx, t4, hw), o in zip(host, outs_h):
            model.lp_sr_host(*t4, prior, hw, out=o)

    step_host(); step_host()
    _barrier(dist, world, dev)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    _barrier(dist, world, dev)
    e2e_ms = _max_over_ranks(dist, world, dev, (time.perf_counter() - t0) * 1e3) / args.steps
    h2d = sum(sum(t.numel() for t in t4) * 4 for s_, idx, t4, hw in host) * world
    d2h = sum(o.numel() * 4 for o in outs_h) * world
    # ---- roofline of the conv class (all tcgen05 launches of one step)
    peaks = load_peaks()
    L.bfsr_prof_enable(1)
    step()
    tms, work, cnt = C.c_double(), C.c_double(), C.c_int64()
    L.bfsr_prof_summary(1, C.byref(tms), C.byref(work), C.byref(cnt))
    oms, owork, ocnt = C.c_double(), C.c_double(), C.c_int64()
    L.bfsr_prof_summary(3, C.byref(oms), C.byref(owork), C.byref(ocnt))
    L.bfsr_prof_enable(0)
    ach = work.value / (tms.value * 1e-3) / 1e12 if tms.value else 0.0
    roofline = {"bound": "tensor", "kernel": "conv_tcgen05", "achieved": ach, "peak": peaks["tflops"], "unit": "TFLOP/s",
                "frac": ach / peaks["tflops"], "traffic": None, "peak_source": peaks["src"], "launches": cnt.value,
                "kernel_ms_per_step": tms.value, "share_of_step": tms.value / ms if ms else None,
                "peak_3pass": peaks["tflops"] / 3.0, "frac_of_3pass_peak": ach / (peaks["tflops"] / 3.0),
                "other_kernels_ms_per_step": oms.value,
                "note": "all tcgen05 conv launches of one step (encoder, coef/freq, MLP, prior): algorithmic conv FLOPs / summed CUDA-event "
                        "time; the query-side kernels (Fourier features, 27-d flow both directions) are in other_kernels_ms_per_step"}
    # ---- CPU baseline + parity on a bounded sample (rank 0, N = 1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import linf_oracle as LO
        torch.set_num_threads(os.cpu_count() or 1)
        s_, idx, ins = buckets[0]
        n = min(2, len(idx))
        c_in = [t[:n].cpu() for t in ins[:4]]
        LO.lp_sr(sd, psd, enc, *[t[:1] for t in c_in], ins[4], literal=True)
        t0 = time.perf_counter()
        ref = LO.lp_sr(sd, psd, enc, *c_in, ins[4], literal=True)
        dt = time.perf_counter() - t0
        got = model.lp_sr(*[t[:n] for t in ins[:4]], prior, ins[4]).cpu().double()
        cpu = {"value": n * ins[4][0] * ins[4][1] / dt / 1e6, "unit": "HR-Mpix/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{n} of {B} patches at scale {s_}, literal reference work (encoder and MLP in both passes), {dt:.2f} s",
               "parity_rel_l2_vs_gpu": float((got - ref.double()).norm() / ref.double().norm())}
    if rank == 0:
        wl = (f"LINF-LP EDSR-baseline 4x, {B} patches of 48x48 LR (q = 65), {wsrc}" if cfg == 3 else
              f"LINF-LP RRDB, {B} patches of 48x48 LR, scales 2/3/4/6/8 cycling, buckets sharded over {world} GPU(s), {wsrc}")
        print(json.dumps({
            "metric": f"HR Mpixels/sec, LINF-LP LP inference (BASELINE config {cfg})", "value": value, "unit": "HR-Mpix/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if args.precision == 0 else "bf16", "data": "synthetic",
            "config": {"workload": wl, "global_batch": B, "precision_mode": "fp32-accurate (split-bf16 x3 on tcgen05)" if args.precision == 0 else "bf16-fast",
                       "l2": "256 MB flush buffer written between timed steps", "cuda_graph": os.environ.get("BFSR_GRAPH", "1") != "0",
                       "parallelism": f"dp{world} (independent patches, no data-path collective)"},
            "clocks": cs.summary(), "e2e": {"value": hr_px_job / (e2e_ms * 1e-3) / 1e6, "unit": "HR-Mpix/s", "h2d_bytes_per_step": h2d,
                                            "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_config4(args, rank, world, local):
    """Config 4: SRFlow-LP 8x (RRDB nb=23, K=16, L=4, two Split2d, three latents) on 80x80 LR tiles, global batch 64 sharded over
    the ranks (16 tiles per GPU at N = 4), with the three-branch generalisation of the SRFlow-LP prior (unet.py:109-181, see
    DESIGN.md).  Synthetic weights; the prior's output convs are zero so the untrained 64-step inverse stays finite (any non-zero
    latent overflows it: oracle/make_golden_r2.py) -- the launch sequence and the arithmetic per launch do not depend on the data."""
    import torch.distributed as dist
    from bfsr_b200 import _lib, models
    from bfsr_b200.dist import shard_range
    from tools import synth
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        saved_fd = os.dup(1); os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev); dist.barrier(); torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush(); os.dup2(saved_fd, 1); os.close(saved_fd)
    B = args.batch if args.batch != BATCH else 64
    S8, LR = 8, 80
    t = synth.SRFlowTopo(scale=8, L=4)
    net = models.define_Flow(t.opt(), device=dev, tile_chunk=args.tile_chunk, precision=args.precision)
    net.load_state_dict(synth.synth_srflow_state_dict(t, seed=31), strict=True)
    lat_ch = (6, 12, 192)
    usd = synth.synth_unet_state_dict(synth.unet_srflow_param_shapes(latent_ch=lat_ch), seed=32)
    for k in usd:
        if k.startswith("outc"):
            usd[k] = torch.zeros_like(usd[k])
    prior = models.make({"name": "unet", "args": {"depth": 3, "dim": 64, "bilinear": True, "latent_ch": lat_ch}, "sd": usd}, load_sd=True)
    lo, hi = shard_range(B, world, rank)
    lr_host = synth.img(B, LR, LR, 1238)[lo:hi].contiguous().pin_memory()
    lr = lr_host.to(dev)
    L = _lib.lib()
    warm = max(args.warmup, 3)
    # one output buffer for every call, as in the config-2 loop: the engine replays its captured launch plan only while the buffers
    # stay the same (a fresh torch.empty per call alternates between two addresses = two plans)
    sr_buf = torch.empty((hi - lo, 3, S8 * LR, S8 * LR), device=dev, dtype=torch.float32)
    for _ in range(warm):
        sr = net.lp_sr(lr, prior, out=sr_buf)
    _barrier(dist, world, dev)
    L.bfsr_launch_count(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as cs:
        _barrier(dist, world, dev)
        e0.record()
        for _ in range(args.steps):
            sr = net.lp_sr(lr, prior, out=sr_buf)
        e1.record()
        _barrier(dist, world, dev)
    launches = int(L.bfsr_launch_count(0))
    ms = _max_over_ranks(dist, world, dev, e0.elapsed_time(e1)) / args.steps
    hr_px_job = B * (S8 * LR) ** 2
    value = hr_px_job / (ms * 1e-3) / 1e6
    assert torch.isfinite(sr).all()
    out_host = torch.empty((hi - lo, 3, S8 * LR, S8 * LR), dtype=torch.float32).pin_memory()
    net.lp_sr_host(lr_host, prior, out=out_host)
    _barrier(dist, world, dev)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        net.lp_sr_host(lr_host, prior, out=out_host)
    _barrier(dist, world, dev)
    e2e_ms = _max_over_ranks(dist, world, dev, (time.perf_counter() - t0) * 1e3) / args.steps
    peaks = load_peaks()
    # ~183 M MAC per LR pixel (flow + encoder 165.14 M, three-branch prior ~18.1 M, estimated)
    alg_tflop = 2 * 183.2e6 * B * LR * LR / 1e12
    if rank == 0:
        print(json.dumps({
            "metric": "HR Mpixels/sec, SRFlow-LP 8x LP inference (BASELINE config 4)", "value": value, "unit": "HR-Mpix/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if args.precision == 0 else "bf16", "data": "synthetic",
            "config": {"workload": f"SRFlow-LP 8x RRDB(nb=23,K=16,L=4) {LR}x{LR} LR tiles, global batch {B} ({hi - lo} per GPU), three-branch prior, synthetic weights",
                       "global_batch": B, "tiles_per_gpu": hi - lo, "l2": "per-step working set (tens of GB) >> 126 MB L2, no flush needed",
                       "parallelism": f"dp{world} (independent tiles, no data-path collective)"},
            "clocks": cs.summary(), "e2e": {"value": hr_px_job / (e2e_ms * 1e-3) / 1e6, "unit": "HR-Mpix/s", "h2d_bytes_per_step": lr_host.numel() * 4 * world,
                                            "d2h_bytes_per_step": out_host.numel() * 4 * world, "ms_per_step": e2e_ms},
            "gpu_launches": launches, "alg_tflop_per_step": alg_tflop,
            "path_tensor_roofline_frac": (alg_tflop / world / (ms * 1e-3)) / peaks["tflops"], "roofline": None, "cpu_baseline": None}), flush=True)
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="BASELINE.json config: 2 = SRFlow-LP 4x (headline, default), 3 = LINF-LP EDSR 48x48 b64, 4 = SRFlow-LP 8x 80x80 b64, 5 = LINF-LP RRDB mixed scales b256")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.config in (3, 5):
        run_linf(args, rank, world, local, args.config)
        return
    if args.config == 4:
        run_config4(args, rank, world, local)
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
        buf = torch.empty((x.shape[0], 3, SCALE * S, SCALE * S), device=dev, dtype=torch.float32)
        for _ in range(2):
            net.lp_sr(x, prior, out=buf)           # same buffers from here on: the plan is captured once, then replayed
        barrier()
        L.bfsr_launch_count(1)                     # count the launches of the timed steps only
        e0.record()
        for _ in range(steps):
            out = net.lp_sr(x, prior, out=buf)
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
    dump = C.create_string_buffer(1 << 20)
    L.bfsr_prof_dump(dump, len(dump))
    L.bfsr_prof_enable(0)
    # one-launch coupling steps (coupling_fused.cu): HBM-side view -- compulsory bytes per level pixel (pre-activation 256 + z 48/96 in
    # and out + hF 96/192 + z1 operand 32/64 in and out) x pixels / device time, against the measured copy bandwidth
    cpl_roof = None
    for ln in dump.value.decode().splitlines():
        tag, n, tms_, _work = ln.split("\t")
        if tag.startswith("cpl-fused C12") and float(tms_) > 0 and not args.tile_chunk:
            hw = tag.split()[-1].split("x")
            px = B_local * int(hw[0]) * int(hw[1])
            byt = (256 + 48 + 96 + 32 + 48 + 32) * px * int(n)
            ach = byt / (float(tms_) * 1e-3) / 1e9
            cpl_roof = {"bound": "hbm", "kernel": "coupling_fused_kernel<12> (" + tag + ")", "achieved": ach, "peak": peaks["hbm_gbs"],
                        "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "launches": int(n), "kernel_ms_per_step": float(tms_),
                        "algorithmic_bytes_per_pixel": 512}
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
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "roofline_flowstep": flow_roof, "roofline_coupling": cpl_roof,
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
