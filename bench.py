#!/usr/bin/env python
"""Benchmark of the hot path: samples/s of (1000-step class-conditional DDPM in latent space + VAE decode).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--precision bf16|fp32]
                    [--workload v2|v3|v4] [--no-cpu] [--no-secondary]

One "step" = one pass of the hot path over one batch: x_T draw, the 1000-step reverse loop (ONE persistent kernel
launch) and the decoder, for `--batch` samples per GPU (default 256 = BASELINE.json configs[1]).  Under torchrun
(N > 1) every rank runs its own shard (weak scaling: 256 samples per GPU, 2048 in total at N = 8) with no per-step
communication; the decoded images are all-gathered once per step with NCCL and the time is the max over ranks.
Prints ONE JSON line (rank 0).  Besides the base contract the v2 line carries

  roofline          the dominant kernel (chain_kernel), CUDA-event timed; `traffic` from the sidecar the profiling
                    script writes (profiles/*chain_traffic.json), never a constant in this file
  roofline_kernels  every kernel of one traced pass (chain + the 38 decoder launches): {name, what, bound, flop | bytes,
                    us, achieved, peak, frac}; tensor-bound ones against the measured bf16 peak, memory-bound ones
                    (norm / gating / LayerNorm passes) as GB/s against the measured HBM copy bandwidth.  Timed with a
                    CUDA event after every launch (ldm_debug_ktrace), not under a profiler
  strong            BASELINE configs[2]: global batch 2048 split over the N GPUs (2048 / N per GPU)
  secondary         configs[3] (v3, 128 rows per GPU = one reference call) and configs[4] (v4, 64 images per GPU): value,
                    ms_per_step, roofline, e2e, clocks; v4 with its own per-kernel list incl. the HBM-bound update

--impl reference times the reference algorithm on the host CPU cores: the oracle port (oracle/restate*.py, pinned
bit-for-bit to the reference in the build container) because /root/reference does not travel to the GPU box.
"""
import argparse
import datetime
import glob
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic work per unit
FLOP_DENOISER_PER_SAMPLE_STEP = 12_845_056          # algorithmically necessary (6 422 528 MAC)
FLOP_DECODE_PER_SAMPLE = 2_809_570_560
N_STEPS = 1000
LATENT = 256
IMG_BYTES = 3 * 64 * 64 * 4
STRONG_GLOBAL_BATCH = 2048                           # BASELINE configs[2]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "src": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "src": "fallback (B200_PROFILING.md)"}


def chain_traffic(batch, precision):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE chain_kernel launch, from the newest sidecar written by
    tools/ncu_traffic.py out of an `ncu --set full` capture of this command (None when there is none for this shape)."""
    best = None
    for p in sorted(glob.glob(os.path.join(ROOT, "profiles", "*chain_traffic.json"))):
        try:
            d = json.load(open(p))
        except Exception:
            continue
        if d.get("batch") == batch and d.get("precision", "bf16") == precision:
            best = (d, os.path.relpath(p, ROOT))
    if best is None:
        return None, "no ncu sidecar for this batch / precision under profiles/ (tools/ncu_traffic.py writes one)"
    return best[0]["dram_bytes_per_launch"], "%s (%s)" % (best[1], best[0].get("source", "ncu --set full"))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons, sampled for the whole run; window(t0, t1) reports the samples that fall
    inside a wall-clock window (the sampler takes a second to come up, so it is started early)."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "25"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def window(self, t0=None, t1=None):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None or not self.path:
            return out
        sm, mx, reasons, every = [], [], set(), []
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                v, m = float(f[2]), float(f[3])
            except ValueError:
                continue
            every.append((v, m))
            if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.05):
                continue
            sm.append(v); mx.append(m)
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[6:10]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm and every:          # window shorter than the sampling period: fall back to the samples under load
            hi = max(v for v, _ in every)
            sm = [v for v, _ in every if v >= 0.5 * hi]; mx = [m for _, m in every]
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out

    def close(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            os.unlink(self.path)
        except OSError:
            pass
        self.proc = None


def clocks_of(sampler, t0, t1):
    c = sampler.window(t0, t1)
    return {"sm_mhz": c["sm_mhz"], "sm_max_mhz": c["sm_max_mhz"], "reasons": c["reasons"]}


# ------------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm on host cores (oracle port), bounded sample
# ------------------------------------------------------------------------------------------------------
def cpu_sample_rate(batch, denoise_steps, decode_rows, threads=None, v3=False):
    """samples/s of the reference path on the CPU, extrapolated from `denoise_steps` of the 1000 reverse steps at the
    full batch and a decode of `decode_rows` samples (both scale linearly: the loop is 1000 identical steps, the
    decoder is per-sample)."""
    import torch
    from oracle import philox, restate as R, weights
    torch.set_grad_enabled(False)
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd_u = weights.make_unet3_state(44, "init") if v3 else weights.make_unet_state(42, "init")
    sd_a = weights.make_decoder_state(43, "init")
    sched = R.schedule(N_STEPS)
    c = torch.arange(batch) % 102
    k = (torch.arange(batch) * 7) % 10
    x = torch.from_numpy(philox.normal_rows(1234, 0, batch, N_STEPS))

    def one(x, t):
        # the reference AS EXECUTED: full multi_head_attention_forward, embeddings recomputed every step, randn_like
        if v3:
            return R.ddpm_update(sched, x, R.unet3_forward(sd_u, x, torch.tensor([t]), c, k), t, torch.randn_like(x) if t > 0 else None)
        return R.p_sample(sd_u, sched, x, t, c, literal_attention=True)

    one(x, N_STEPS - 1)          # warm-up (thread pool, MKL plans)
    t0 = time.perf_counter()
    for t in range(N_STEPS - 1, N_STEPS - 1 - denoise_steps, -1):
        x = one(x, t)
    t_loop = (time.perf_counter() - t0) / denoise_steps * N_STEPS
    z = x[:decode_rows]
    t0 = time.perf_counter()
    R.decode(sd_a, z)
    t_dec = (time.perf_counter() - t0) / decode_rows * batch
    return batch / (t_loop + t_dec), {"loop_s_extrapolated": t_loop, "decode_s_extrapolated": t_dec, "threads": torch.get_num_threads()}


def run_reference_arm(args, rank):
    if rank != 0:
        return
    v3 = args.workload == "v3"
    batch = args.batch
    vals = []
    for i in range(args.warmup + args.steps):
        v, info = cpu_sample_rate(batch, denoise_steps=20, decode_rows=8, v3=v3)
        if i >= args.warmup:
            vals.append(v)
    value = sum(vals) / len(vals)
    sample = "per step: batch %d, 20 of the 1000 reverse steps + decode of 8 samples, extrapolated linearly" % batch
    name = "v3 multi-conditional (class + color) latent U-Net" if v3 else "v2 latent U-Net"
    line = {
        "impl": "reference", "metric": "samples/s (1000-step DDPM + VAE decode)", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * batch / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s 1000-step sampling + VAE decode, batch %d per GPU" % (name, batch),
                   "batch_per_gpu": batch, "global_batch": batch, "n_steps": N_STEPS, "precision": "fp32",
                   "where": "host CPU cores (rank 0 only), reference algorithm as executed", "weights": "random-init (seeded), eval mode"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": info["threads"], "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# per-kernel records (roofline_kernels)
# ------------------------------------------------------------------------------------------------------
def decoder_kernel_table(B):
    """The 38 launches of one bf16 decode of B latents, in launch order (csrc/decoder.cu: decode_chunk_bf16):
    (trace name, what, bound, flop or bytes).  Memory-bound passes: algorithmic bytes = tensors read + written once."""
    t = []
    E = lambda C, H: B * H * H * C      # elements of an NHWC activation
    t.append(("launch_load_x", "z fp32 -> bf16", "hbm", B * LATENT * 6))
    t.append(("gemm_tc", "fc.0 256->512", "tensor", 2 * B * 512 * LATENT))
    t.append(("launch_row_ln", "fc.1 LayerNorm(512)+Swish", "hbm", B * 512 * 6))
    t.append(("gemm_tc", "fc.3 512->32768", "tensor", 2 * B * 32768 * 512))
    t.append(("launch_row_ln", "fc.4 LayerNorm(32768)+Swish", "hbm", B * 32768 * 6))
    for C, H in ((512, 8), (256, 16), (128, 32)):
        n = E(C, H)
        conv = 2 * n * 9 * C
        t += [("conv_tc", "res%d.conv1 3x3 %d@%dx%d" % (C, C, H, H), "tensor", conv),
              ("launch_norm_coef_bf16", "res%d.ln1 statistics -> (scale, shift)" % C, "hbm", n * 2),
              ("launch_coef_apply_bf16", "res%d.ln1 apply + Swish" % C, "hbm", n * 4),
              ("conv_tc", "res%d.conv2 3x3" % C, "tensor", conv),
              ("launch_norm_coef_bf16", "res%d.ln2 statistics" % C, "hbm", n * 2),
              ("launch_sa_map_gate", "res%d channel gate + spatial mean/max map + sigmoid(conv7x7(map)), one CTA per sample" % C, "hbm",
               n * 2 + B * H * H * 4),
              ("launch_sa_apply_bf16", "res%d ln2 * CA * SA gate + x, Swish" % C, "hbm", n * 6)]
        no = E(C // 2, 2 * H)
        t += [("conv_tc", "up ConvT(4,2,1) %d->%d @%dx%d" % (C, C // 2, 2 * H, 2 * H), "tensor", 2 * no * 4 * C),
              ("launch_norm_coef_bf16", "up GroupNorm statistics", "hbm", no * 2),
              ("launch_coef_apply_bf16", "up GroupNorm apply + Swish", "hbm", no * 4)]
    n = E(32, 64)
    t += [("conv_halo", "final_conv.0 3x3 64->32 @64x64 (halo kernel)", "tensor", 2 * n * 9 * 64),
          ("launch_norm_coef_bf16", "final GroupNorm(8,32) statistics", "hbm", n * 2),
          ("final_gn_conv3", "final GroupNorm apply + Swish + final_conv.3 3x3 32->3 + Sigmoid (one pass, mma.sync)", "hbm",
           n * 2 + B * 3 * 4096 * 4)]
    return t


def pix_kernel_table(B):
    """The 18 launches of one v4 reverse step at 64 x 64 (csrc/pixel.cu), in launch order: MACs from v4:54-96."""
    macs = [("pix_conv_in", "conv1.0 3->64 (K=27, hi/lo bf16 split)", 4096 * 64 * 27), ("conv_halo", "conv1.2 64->64", 4096 * 64 * 576),
            ("conv_tc", "down1 4x4 s2 64->128", 1024 * 128 * 1024), ("conv_tc", "conv2.0", 1024 * 128 * 1152), ("conv_tc", "conv2.2", 1024 * 128 * 1152),
            ("conv_tc", "down2 4x4 s2 128->256", 256 * 256 * 2048), ("conv_tc", "conv3.0", 256 * 256 * 2304), ("conv_tc", "conv3.2", 256 * 256 * 2304),
            ("conv_tc", "bottleneck.0 256->512", 256 * 512 * 2304), ("conv_tc", "bottleneck.2 512->256", 256 * 256 * 4608),
            ("conv_tc", "up1 ConvT 256->128", 1024 * 128 * 1024), ("conv_tc", "conv4.0 256->128", 1024 * 128 * 2304), ("conv_tc", "conv4.2", 1024 * 128 * 1152),
            ("conv_tc", "up2 ConvT 128->64", 4096 * 64 * 512), ("conv_tc", "conv5.0 128->64", 4096 * 64 * 1152), ("conv_halo", "conv5.2 64->64", 4096 * 64 * 576),
            ("conv_halo", "out_conv 64->3 (N padded to 16)", 4096 * 3 * 576)]
    t = [(n, w, "tensor", 2 * m * B) for n, w, m in macs]
    t.append(("pix_ddpm", "posterior update, Philox in-kernel (12 B / element)", "hbm", B * 3 * 4096 * 12))
    return t


def kernel_records(trace, table, pk, anchor=None):
    """Zip a traced launch list with the expected table; a name mismatch degrades to time-only records."""
    trace = [("conv_tc" if n == "conv_tc_pair" else n, ms) for n, ms in trace]   # the 2-CTA variant of the same convolution
    names = [n for n, _ in trace]
    start = 0
    if anchor is not None:
        start = len(names) - 1 - names[::-1].index(anchor) + 1 if anchor in names else 0
    part = trace[start:start + len(table)]
    ok = len(part) == len(table) and all(p[0] == e[0] for p, e in zip(part, table))
    recs = []
    if not ok:
        return [{"name": n, "us": ms * 1000.0} for n, ms in trace[start:]], False
    for (name, ms), (_, what, bound, work) in zip(part, table):
        s = ms / 1000.0
        if bound == "tensor":
            ach, peak, unit = work / s / 1e12 if s > 0 else None, pk["bf16_sustained"], "TFLOP/s"
            recs.append({"name": name, "what": what, "bound": bound, "flop": work, "us": ms * 1000.0, "achieved": ach, "peak": peak,
                         "unit": unit, "frac": ach / peak if ach else None})
        else:
            ach, peak, unit = work / s / 1e9 if s > 0 else None, pk["hbm_gbs"], "GB/s"
            recs.append({"name": name, "what": what, "bound": bound, "bytes": work, "us": ms * 1000.0, "achieved": ach, "peak": peak,
                         "unit": unit, "frac": ach / peak if ach else None})
    return recs, True


def summarize_hbm(recs):
    """Aggregate GB/s of the memory-bound records of a list (the 'fused elementwise / norm kernels' of north_star)."""
    hb = [r for r in recs if r.get("bound") == "hbm" and r.get("us")]
    if not hb:
        return None
    by, us = sum(r["bytes"] for r in hb), sum(r["us"] for r in hb)
    return {"kernels": len(hb), "bytes": by, "us": us, "achieved_gbs": by / (us * 1e-6) / 1e9}


# ------------------------------------------------------------------------------------------------------
# GPU arm: v2 (headline) and v3 latent workloads
# ------------------------------------------------------------------------------------------------------
V2_NOTE = ("the loop is a chain of 5 dependent contractions per step x 1000 steps inside 16-CTA clusters: each phase is "
           "bounded by one tcgen05.mma of M=128, N<=128 per 67-74 clocks (240 per step and CTA), by the TMA issue rate of its "
           "producer threads and by two cluster-wide exchanges (LayerNorm statistics, operand hand-over), not by the "
           "tensor pipe's peak (DESIGN.md 5.1, tools/ubench.cu); the decoder convolutions are the tensor-bound kernels")
V3_NOTE = ("the rows of a v3 call are coupled by the cross-batch attention, so the call is one grid: 19 dependent phases per step, "
           "each paying a grid-wide barrier (~1.5-2 us: write drain, one atomic, one poll) and ~1.2 us of first-tile latency "
           "before ~1-3 us of work (DESIGN.md 8.1, tools/v3loop_trace.py); the tensor pipe is idle most of the step")


def loop_kernel_name(eng, nb):
    if int(eng.info("chain")):
        return ("chain_kernel: the whole 1000-step loop is ONE persistent launch (16-CTA clusters, 5 dependent tcgen05 "
                "contractions per step, TMA operands, DSMEM statistics)")
    n = int(eng.info("launches_per_step"))
    if n == 0:   # v3: unet3_loop_kernel
        return ("unet3_loop_kernel: the whole 1000-step loop is ONE persistent launch (128 co-resident CTAs, 19 phases per step "
                "separated by grid-wide barriers: folded tcgen05 contractions, LayerNorm rows, one attention head per CTA)")
    return "sampling loop (one CUDA-graph launch = 1000 steps x %d kernels)" % n


def run_latent(workload, B, K, W, prec, rank, world, local_rank, sampler, with_cpu, with_kernels, label):
    """v2 or v3: returns the JSON line (dict) on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist
    import ldm_b200              # nothing under oracle/ is imported on this arm

    torch.set_grad_enabled(False)
    dev = torch.device("cuda", local_rank)
    v3 = workload == "v3"
    # random-init weights of the named architecture (ConditionalUNet + init_weights, SimpleAutoencoder())
    torch.manual_seed(42)                                                       # v2:17
    unet = (ldm_b200.v3 if v3 else ldm_b200).ConditionalUNet(precision=prec)
    unet.apply(ldm_b200.init_weights)                                           # v2:1346-1350 (main()'s initialisation of the denoiser)
    unet = unet.to(dev).eval()
    ae = ldm_b200.SimpleAutoencoder(precision=prec)                             # initialises itself (v2:324)
    ae = ae.to(dev).eval()
    diffusion = (ldm_b200.v3 if v3 else ldm_b200).ConditionalDenoiseDiffusion(unet, N_STEPS, dev)
    eng = unet.engine(dev, N_STEPS)
    eng.set_schedule(*diffusion._host_schedule)
    eng.pack_decoder(ae.decoder)

    total = B * world
    lo = rank * B
    classes = (torch.arange(total) % 102)
    c_dev = classes[lo:lo + B].to(dev)
    k_all = (torch.arange(total) * 7) % 10                                      # v3: colour labels
    k_dev = k_all[lo:lo + B].to(dev)
    gathered = torch.empty(total, 3, 64, 64, device=dev) if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def draw(seed, use_graph=True):
        if v3:
            return diffusion.sample((B, LATENT), dev, c_dev, k_dev, seed=seed, sample_offset=lo, use_graph=use_graph)
        return diffusion.sample((B, LATENT), dev, c_dev, seed=seed, sample_offset=lo, use_graph=use_graph)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        img = ae.decode(draw(1234 + i))
        if world > 1:
            dist.all_gather_into_tensor(gathered, img)
    barrier()
    eng.check_device_flags()

    # ---- device-timed region: K steps, L2 flushed between iterations (flush outside the event pairs)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(K)]
    launches0 = eng.launches()
    barrier()
    t_win0 = time.time()
    for i in range(K):
        flush.zero_()
        s, m, e = ev[i]
        s.record()
        x0 = draw(99 + i)
        m.record()
        img = ae.decode(x0)
        if world > 1:
            dist.all_gather_into_tensor(gathered, img)
        e.record()
    barrier()
    launches = eng.launches() - launches0
    t_loop_ms = sum(s.elapsed_time(m) for s, m, e in ev)
    t_total_ms = sum(s.elapsed_time(e) for s, m, e in ev)
    assert torch.isfinite(img).all()

    # ---- end to end through the host-buffer C entry point: pinned labels in, pinned images out, every step
    c_host = classes[lo:lo + B].clone().pin_memory()
    k_host = k_all[lo:lo + B].clone().pin_memory()
    img_host = torch.empty(B, 3, 64, 64).pin_memory()

    def host_call(seed):          # synchronous: returns with the images on the host
        if v3:
            eng.generate3_host(c_host, k_host, img_host, None, seed=seed, sample_offset=lo)
        else:
            eng.generate_host(c_host, img_host, None, seed=seed, sample_offset=lo)
    host_call(5)
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        host_call(500 + i)
    barrier()
    t_e2e = time.perf_counter() - t0
    t_win1 = time.time()

    # ---- one traced pass (CUDA event after every launch) for the per-kernel records; outside every timed region
    trace = None
    if with_kernels and rank == 0:
        eng.ktrace_start()
        ae.decode(draw(7, use_graph=False))
        trace = eng.ktrace_stop()

    if world > 1:
        t = torch.tensor([t_total_ms, t_loop_ms, t_e2e * 1000.0], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_total_ms, t_loop_ms, t_e2e = float(t[0]), float(t[1]), float(t[2]) / 1000.0
    del flush, gathered
    if rank != 0:
        return None
    pk = peaks()
    value = total * K / (t_total_ms / 1000.0)
    flop_step = FLOP_DENOISER_PER_SAMPLE_STEP if not v3 else 2 * (9_633_792 + 2 * B * 2304)   # v3: + Q, K projections and B x B attention
    loop_flops = flop_step * B * N_STEPS                                          # per launch, per GPU
    loop_s = t_loop_ms / 1000.0 / K
    achieved = loop_flops / loop_s / 1e12
    dec_s = (t_total_ms - t_loop_ms) / 1000.0 / K
    chain = bool(int(eng.info("chain")))
    traffic, traffic_src = chain_traffic(B, prec) if (chain and not v3) else (None, "not captured for this workload")
    line = {
        "metric": "samples/s (1000-step DDPM + VAE decode)", "value": value, "unit": "samples/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": t_total_ms / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if prec == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": ("v3 multi-conditional (class + color) latent U-Net" if v3 else "v2 latent U-Net") +
                               " 1000-step sampling + VAE decode, batch %d per GPU%s%s" %
                               (B, "" if world == 1 else ", %d total, final NCCL all-gather of images" % total, label),
                   "batch_per_gpu": B, "global_batch": total, "n_steps": N_STEPS, "precision": prec,
                   "l2": "flushed (256 MiB write) between timed iterations", "weights": "random-init (seeded), eval mode"},
        "e2e": {"value": total * K / t_e2e, "unit": "samples/s", "h2d_bytes_per_step": B * (16 if v3 else 8), "d2h_bytes_per_step": B * IMG_BYTES},
        "gpu_launches": int(launches),
        "ms_per_iter": [round(s.elapsed_time(e), 3) for s, m, e in ev],      # this rank's timed iterations (CUDA events)
        "clocks": clocks_of(sampler, t_win0, t_win1),
        "roofline": {"bound": "tensor", "kernel": loop_kernel_name(eng, B),
                     "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_sustained"],
                     "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": pk["src"], "ms_per_launch": loop_s * 1000.0,
                     "algorithmic_flop_per_launch": loop_flops,
                     "note": V3_NOTE if v3 else V2_NOTE},
        "decode": {"ms": dec_s * 1000.0, "tflops": FLOP_DECODE_PER_SAMPLE * B / dec_s / 1e12 if dec_s > 0 else None},
    }
    if trace is not None and prec == "bf16":
        recs, ok = [], True
        ck = [(n, ms) for n, ms in trace if n == "chain_kernel"]
        if ck:
            s = ck[-1][1] / 1000.0
            recs.append({"name": "chain_kernel", "what": "1000 reverse steps, fused posterior update", "bound": "tensor", "flop": loop_flops,
                         "us": ck[-1][1] * 1000.0, "achieved": loop_flops / s / 1e12, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                         "frac": loop_flops / s / 1e12 / pk["bf16_sustained"]})
        dec_trace = trace[len(trace) - len(decoder_kernel_table(B)):] if len(trace) >= len(decoder_kernel_table(B)) else trace
        drecs, ok = kernel_records(dec_trace, decoder_kernel_table(B), pk)
        line["roofline_kernels"] = recs + drecs
        line["roofline_kernels_info"] = {"timing": "CUDA event after every launch of one untimed eager pass (ldm_debug_ktrace); us includes the gap to the previous kernel",
                                         "table_matched": ok, "hbm_peak_gbs": pk["hbm_gbs"], "hbm_kernels": summarize_hbm(drecs) if ok else None,
                                         "decoder_us_traced": sum(r["us"] for r in drecs)}
    if world == 1 and with_cpu:
        v, info = cpu_sample_rate(B, denoise_steps=60 if not v3 else 30, decode_rows=32, v3=v3)
        line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": info["threads"], "kind": "port",
                                "sample": "batch %d: %d of the 1000 reverse steps + decode of 32 samples, extrapolated linearly" % (B, 60 if not v3 else 30)}
    return line


# ------------------------------------------------------------------------------------------------------
# v4 pixel-space workload (BASELINE configs[4])
# ------------------------------------------------------------------------------------------------------
PIX_MAC_PER_SAMPLE_STEP = (4096 * 64 * 27 + 4096 * 64 * 576 + 1024 * 128 * 1024 + 2 * 1024 * 128 * 1152 + 256 * 256 * 2048
                           + 2 * 256 * 256 * 2304 + 256 * 512 * 2304 + 256 * 256 * 4608 + 1024 * 128 * 1024 + 1024 * 128 * 2304
                           + 1024 * 128 * 1152 + 4096 * 64 * 512 + 4096 * 64 * 1152 + 4096 * 64 * 576 + 4096 * 3 * 576)   # v4:54-96 at 64 x 64


def pix_cpu_rate(batch, steps, threads=None):
    """images/s of the v4 reference path (oracle/restate_pix.py, bit-equal to the reference) on the host cores,
    extrapolated from `steps` of the 1000 reverse steps at `batch` images."""
    import torch
    from oracle import philox, restate as R, restate_pix as P, weights
    torch.set_grad_enabled(False)
    torch.set_num_threads(threads or os.cpu_count() or 1)
    sd = weights.make_pix_state(45, "init")
    sched = R.schedule(N_STEPS)
    x = torch.from_numpy(philox.normal_rows(1234, 0, batch, N_STEPS, 3 * 64 * 64)).view(batch, 3, 64, 64)
    P.p_sample(sd, sched, x, N_STEPS - 1)
    t0 = time.perf_counter()
    for t in range(N_STEPS - 1, N_STEPS - 1 - steps, -1):
        x = P.p_sample(sd, sched, x, t)
    dt = (time.perf_counter() - t0) / steps * N_STEPS
    return batch / dt, torch.get_num_threads()


def run_pix(B, K, W, rank, world, local_rank, sampler, with_cpu, with_kernels):
    import torch
    import torch.distributed as dist
    from ldm_b200 import v4      # nothing under oracle/ is imported on this arm

    torch.set_grad_enabled(False)
    dev = torch.device("cuda", local_rank)
    torch.manual_seed(42)
    model = v4.SimpleUNet()                                                    # torch's default initialisation (v4 never re-initialises)
    model = model.to(dev).eval()
    diffusion = v4.DiffusionModel(model, N_STEPS, device=dev)
    eng = diffusion._engine(dev)
    total, lo = B * world, rank * B
    gathered = torch.empty(total, 3, 64, 64, device=dev) if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(seed):
        img = diffusion.sample((B, 3, 64, 64), seed=seed, sample_offset=lo)
        if world > 1:
            dist.all_gather_into_tensor(gathered, img)
        return img

    for i in range(W):
        step(1234 + i)
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    launches0 = eng.launches()
    barrier()
    t_win0 = time.time()
    for i in range(K):
        flush.zero_()
        ev[i][0].record()
        img = step(99 + i)
        ev[i][1].record()
    barrier()
    launches = eng.launches() - launches0
    t_ms = sum(s.elapsed_time(e) for s, e in ev)
    assert torch.isfinite(img).all()
    # end to end: images land in pinned host memory every step (the only per-step input is the seed)
    img_host = torch.empty(B, 3, 64, 64).pin_memory()
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        img_host.copy_(diffusion.sample((B, 3, 64, 64), seed=500 + i, sample_offset=lo), non_blocking=True)
        torch.cuda.synchronize()
    barrier()
    t_e2e = time.perf_counter() - t0
    t_win1 = time.time()
    trace = None
    if with_kernels and rank == 0:
        x = eng.randn(B, 3 * 64 * 64, 3, lo, N_STEPS).view(B, 3, 64, 64)
        eng.pix_sample(x, 999, 999, seed=3, sample_offset=lo, use_graph=False)        # warm (TMA descriptors of this batch)
        eng.ktrace_start()
        eng.pix_sample(x, 998, 997, seed=3, sample_offset=lo, use_graph=False)        # two eager steps; the second one is reported
        trace = eng.ktrace_stop()
    if world > 1:
        t = torch.tensor([t_ms, t_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_ms, t_e2e = float(t[0]), float(t[1])
    del flush, gathered
    if rank != 0:
        return None
    pk = peaks()
    flop = 2.0 * PIX_MAC_PER_SAMPLE_STEP * B * N_STEPS
    achieved = flop / (t_ms / 1000.0 / K) / 1e12
    line = {
        "metric": "samples/s (1000-step pixel-space DDPM, 64x64 images)", "value": total * K / (t_ms / 1000.0), "unit": "samples/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": t_ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "v4 pixel-space diffusion U-Net 1000-step sampling at 64x64, batch %d per GPU%s" %
                               (B, "" if world == 1 else ", %d total, final NCCL all-gather of images" % total),
                   "batch_per_gpu": B, "global_batch": total, "n_steps": N_STEPS, "precision": "bf16",
                   "l2": "flushed (256 MiB write) between timed iterations", "weights": "random-init (seeded, torch default statistics), eval mode"},
        "e2e": {"value": total * K / t_e2e, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": B * IMG_BYTES},
        "gpu_launches": int(launches),
        "clocks": clocks_of(sampler, t_win0, t_win1),
        "roofline": {"bound": "tensor", "kernel": "sampling loop (one CUDA-graph launch = 1000 steps x %d kernels: 15 tcgen05 implicit-GEMM convolutions, "
                                                  "conv1.0 / out_conv / posterior update)" % (launches // (K * N_STEPS)),
                     "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_sustained"],
                     "traffic": None, "peak_source": pk["src"], "ms_per_launch": t_ms / K,
                     "algorithmic_flop_per_launch": flop},
    }
    if trace is not None:
        tab = pix_kernel_table(B)
        recs, ok = kernel_records(trace[len(trace) - len(tab):], tab, pk)
        line["roofline_kernels"] = recs
        line["roofline_kernels_info"] = {"timing": "CUDA event after every launch of one eager reverse step (ldm_debug_ktrace)", "table_matched": ok,
                                         "hbm_peak_gbs": pk["hbm_gbs"], "hbm_kernels": summarize_hbm(recs) if ok else None,
                                         "step_us_traced": sum(r["us"] for r in recs)}
    if with_cpu:
        v, th = pix_cpu_rate(4, 3)
        line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": th, "kind": "port",
                                "sample": "batch 4, 3 of the 1000 reverse steps, extrapolated linearly (oracle/restate_pix.py, bit-equal to the v4 reference)"}
    return line


def run_pix_reference_arm(args, rank):
    if rank != 0:
        return
    vals = []
    for i in range(args.warmup + args.steps):
        v, th = pix_cpu_rate(min(args.batch, 8), 2)
        if i >= args.warmup:
            vals.append(v)
    value = sum(vals) / len(vals)
    print(json.dumps({
        "impl": "reference", "metric": "samples/s (1000-step pixel-space DDPM, 64x64 images)", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * args.batch / value, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "v4 pixel-space diffusion U-Net 1000-step sampling at 64x64, batch %d per GPU" % args.batch,
                   "where": "host CPU cores (rank 0 only), reference algorithm", "n_steps": N_STEPS},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": th, "kind": "port",
                         "sample": "per step: batch %d, 2 of the 1000 reverse steps, extrapolated linearly" % min(args.batch, 8)},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


def compact(line):
    """The fields of a secondary / strong record."""
    keep = ("metric", "value", "unit", "ms_per_step", "steps", "warmup", "scaling", "dtype", "config", "e2e", "gpu_launches", "clocks", "roofline",
            "roofline_kernels", "roofline_kernels_info", "decode")
    return {k: line[k] for k in keep if k in line}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=None, help="samples per GPU per step (v2: 256, v3: 128, v4: 64)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-secondary", action="store_true", help="v2 only: skip the strong-scaling and v3 / v4 records")
    ap.add_argument("--workload", default="v2", choices=["v2", "v3", "v4"],
                    help="v2: BASELINE configs[1]/[2] (default). v3: configs[3], multi-conditional denoiser, --batch rows per GPU (default 128), "
                         "each GPU one reference call. v4: configs[4], pixel-space U-Net at 64x64, --batch images per GPU (default 64)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.batch is None:
        args.batch = {"v2": 256, "v3": 128, "v4": 64}[args.workload]
    if args.impl == "reference":
        if args.workload == "v4":
            run_pix_reference_arm(args, rank)
        else:
            run_reference_arm(args, rank)
        return
    import torch
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sampler = ClockSampler(local_rank)
    sampler.start()
    try:
        K, W = args.steps, args.warmup
        with_cpu = not args.no_cpu and world == 1
        if args.workload == "v4":
            line = run_pix(args.batch, K, W, rank, world, local_rank, sampler, with_cpu, True)
        else:
            line = run_latent(args.workload, args.batch, K, W, args.precision, rank, world, local_rank, sampler, with_cpu, True, "")
        if args.workload == "v2" and not args.no_secondary and args.precision == "bf16":
            k2, w2 = min(K, 2), 3
            # BASELINE configs[2]: the SAME global batch of 2048 at every N (2048 / N samples per GPU)
            if STRONG_GLOBAL_BATCH % world == 0:
                s = run_latent("v2", STRONG_GLOBAL_BATCH // world, k2, w2, args.precision, rank, world, local_rank, sampler, False, False,
                               " (configs[2]: global batch 2048 strong-split)")
                if rank == 0:
                    s = compact(s)
                    s["scaling"] = "strong"
                    s["global_batch"] = STRONG_GLOBAL_BATCH
                    s["batch_per_gpu"] = STRONG_GLOBAL_BATCH // world
                    line["strong"] = s
            sec3 = run_latent("v3", 128, k2, w2, "bf16", rank, world, local_rank, sampler, False, False, " (configs[3]: 128 rows per GPU = one reference call)")
            sec4 = run_pix(64, k2, w2, rank, world, local_rank, sampler, False, True)
            if rank == 0:
                line["secondary"] = {"v3": compact(sec3), "v4": compact(sec4)}
        if rank == 0:
            print(json.dumps(line))
    finally:
        sampler.close()
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
