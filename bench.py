#!/usr/bin/env python
"""Benchmark of the hot path: samples/s of (1000-step class-conditional DDPM in latent space + VAE decode).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--precision bf16|fp32]

One "step" = one pass of the hot path over one batch: x_T draw, the 1000-step reverse loop (one CUDA-graph
launch) and the decoder, for `--batch` samples per GPU (default 256 = BASELINE.json configs[1]).  Under
torchrun (N > 1) every rank runs its own shard (weak scaling: 256 samples per GPU, 2048 in total at N = 8 =
configs[2]) with no per-step communication; the decoded images are all-gathered once per step with NCCL and
the time is the max over ranks.  Prints ONE JSON line (rank 0).

--impl reference times the reference algorithm on the host CPU cores: the oracle port (oracle/restate.py, pinned
bit-for-bit to the reference in the build container) because /root/reference does not travel to the GPU box.
"""
import argparse
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
CHAIN_DRAM_BYTES_PER_LAUNCH = 33_692_160 + 111_360   # ncu capture r01_g (see roofline.traffic_source)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "src": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons.  The sampler is started early (nvidia-smi takes a second to come up) and
    every line carries a timestamp; stop(t0, t1) keeps the samples that fall inside the wall-clock window of the timed
    region."""
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

    def stop(self, t0=None, t1=None):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
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
        os.unlink(self.path)
        if not sm and every:          # window shorter than the sampling period: fall back to the samples under load
            hi = max(v for v, _ in every)
            sm = [v for v, _ in every if v >= 0.5 * hi]; mx = [m for _, m in every]
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm on host cores (oracle port), bounded sample
# ------------------------------------------------------------------------------------------------------
def cpu_sample_rate(batch, denoise_steps, decode_rows, threads=None):
    """samples/s of the reference path on the CPU, extrapolated from `denoise_steps` of the 1000 reverse steps at the
    full batch and a decode of `decode_rows` samples (both scale linearly: the loop is 1000 identical steps, the
    decoder is per-sample)."""
    import torch
    from oracle import philox, restate as R, weights
    torch.set_grad_enabled(False)
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd_u = weights.make_unet_state(42, "init")
    sd_a = weights.make_decoder_state(43, "init")
    sched = R.schedule(N_STEPS)
    c = torch.arange(batch) % 102
    x = torch.from_numpy(philox.normal_rows(1234, 0, batch, N_STEPS))
    R.p_sample(sd_u, sched, x, N_STEPS - 1, c, literal_attention=True)          # warm-up (thread pool, MKL plans)
    t0 = time.perf_counter()
    for t in range(N_STEPS - 1, N_STEPS - 1 - denoise_steps, -1):
        # the reference AS EXECUTED: full multi_head_attention_forward, embeddings recomputed every step, randn_like
        x = R.p_sample(sd_u, sched, x, t, c, literal_attention=True)
    t_loop = (time.perf_counter() - t0) / denoise_steps * N_STEPS
    z = x[:decode_rows]
    t0 = time.perf_counter()
    R.decode(sd_a, z)
    t_dec = (time.perf_counter() - t0) / decode_rows * batch
    return batch / (t_loop + t_dec), {"loop_s_extrapolated": t_loop, "decode_s_extrapolated": t_dec, "threads": torch.get_num_threads()}


def run_reference_arm(args, rank):
    if rank != 0:
        return
    batch = args.batch
    vals = []
    for i in range(args.warmup + args.steps):
        v, info = cpu_sample_rate(batch, denoise_steps=20, decode_rows=8)
        if i >= args.warmup:
            vals.append(v)
    value = sum(vals) / len(vals)
    sample = "per step: batch %d, 20 of the 1000 reverse steps + decode of 8 samples, extrapolated linearly" % batch
    line = {
        "impl": "reference", "metric": "samples/s (1000-step DDPM + VAE decode)", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * batch / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "v2 latent U-Net 1000-step sampling + VAE decode, batch %d per GPU" % batch,
                   "batch_per_gpu": batch, "global_batch": batch, "n_steps": N_STEPS, "precision": "fp32",
                   "where": "host CPU cores (rank 0 only), reference algorithm as executed", "weights": "random-init (seeded), eval mode"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": info["threads"], "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def loop_kernel_name(eng):
    if int(eng.info("chain")):
        return "chain_kernel: the whole 1000-step loop is ONE persistent launch (16-CTA clusters x 48 samples, 5 dependent tcgen05 contractions per step, TMA operands, DSMEM statistics)"
    return "sampling loop (one CUDA-graph launch = 1000 steps x %d kernels)" % int(eng.info("launches_per_step"))


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import ldm_b200              # nothing under oracle/ is imported on this arm

    torch.set_grad_enabled(False)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    sampler = ClockSampler(local_rank)
    sampler.start()
    B, K, W = args.batch, args.steps, args.warmup
    prec = args.precision

    # random-init weights of the named architecture (ConditionalUNet + init_weights, SimpleAutoencoder())
    v3 = args.workload == "v3"
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
    gathered = torch.empty(total, 3, 64, 64, device=dev) if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)       # > 126 MB L2

    k_dev = ((torch.arange(total) * 7) % 10)[lo:lo + B].to(dev)     # v3: colour labels

    def draw(seed):
        if v3:
            return diffusion.sample((B, LATENT), dev, c_dev, k_dev, seed=seed, sample_offset=lo)
        return diffusion.sample((B, LATENT), dev, c_dev, seed=seed, sample_offset=lo)

    def step(i):
        x0 = draw(1234 + i)
        img = ae.decode(x0)
        if world > 1:
            dist.all_gather_into_tensor(gathered, img)
        return img

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        step(i)
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
    t_win1 = time.time()
    launches = eng.launches() - launches0
    t_loop_ms = sum(s.elapsed_time(m) for s, m, e in ev)
    t_total_ms = sum(s.elapsed_time(e) for s, m, e in ev)
    assert torch.isfinite(img).all()

    # ---- end to end through the host-buffer entry point: pinned labels in, pinned images out, every step
    c_host = classes[lo:lo + B].clone().pin_memory()
    img_host = torch.empty(B, 3, 64, 64).pin_memory()
    if v3:
        k_host = k_dev.cpu().pin_memory()

        def host_call(seed):   # v3 has no single C entry point with host buffers yet: pinned labels in, pinned images out
            f_d, k_d = c_host.to(dev, non_blocking=True), k_host.to(dev, non_blocking=True)
            z = diffusion.sample((B, LATENT), dev, f_d, k_d, seed=seed, sample_offset=lo)
            img_host.copy_(ae.decode(z), non_blocking=True)
            torch.cuda.synchronize()
    else:
        def host_call(seed):
            eng.generate_host(c_host, img_host, None, seed=seed, sample_offset=lo)   # synchronous: returns with the images on the host
    host_call(5)
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        host_call(500 + i)
    barrier()
    t_e2e = time.perf_counter() - t0
    clocks = sampler.stop(t_win0, time.time())   # clocks during the device-timed and the end-to-end regions

    if world > 1:
        t = torch.tensor([t_total_ms, t_loop_ms, t_e2e * 1000.0], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_total_ms, t_loop_ms, t_e2e = float(t[0]), float(t[1]), float(t[2]) / 1000.0

    if rank != 0:
        return
    pk = peaks()
    value = total * K / (t_total_ms / 1000.0)
    flop_step = FLOP_DENOISER_PER_SAMPLE_STEP if not v3 else 2 * (9_633_792 + 2 * B * 2304)   # v3: + Q, K projections and B x B attention
    loop_flops = flop_step * B * N_STEPS                                          # per graph launch, per GPU
    loop_s = t_loop_ms / 1000.0 / K
    achieved = loop_flops / loop_s / 1e12
    dec_s = (t_total_ms - t_loop_ms) / 1000.0 / K
    line = {
        "metric": "samples/s (1000-step DDPM + VAE decode)", "value": value, "unit": "samples/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": t_total_ms / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if prec == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": ("v3 multi-conditional (class + color) latent U-Net" if v3 else "v2 latent U-Net") +
                               " 1000-step sampling + VAE decode, batch %d per GPU%s" %
                               (B, "" if world == 1 else ", %d total, final NCCL all-gather of images" % total),
                   "batch_per_gpu": B, "global_batch": total, "n_steps": N_STEPS, "precision": prec,
                   "l2": "flushed (256 MiB write) between timed iterations", "weights": "random-init (seeded), eval mode"},
        "e2e": {"value": total * K / t_e2e, "unit": "samples/s", "h2d_bytes_per_step": B * (16 if v3 else 8), "d2h_bytes_per_step": B * IMG_BYTES},
        "gpu_launches": int(launches),
        "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"]},
        "roofline": {"bound": "tensor", "kernel": loop_kernel_name(eng),
                     "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_sustained"],
                     "traffic": CHAIN_DRAM_BYTES_PER_LAUNCH if (prec == "bf16" and B == 256 and not v3 and int(eng.info("chain"))) else None,
                     "traffic_source": "ncu --set full, profiles/r01_g_chain_full_summary.txt (dram__bytes_read.sum + dram__bytes_write.sum of one "
                                       "1000-step launch at B=256; weights and operands stay in L2, hit rate 96.7 %)",
                     "peak_source": pk["src"], "ms_per_launch": loop_s * 1000.0,
                     "algorithmic_flop_per_launch": loop_flops,
                     "note": "the loop is a chain of 5 dependent contractions per step x 1000 steps on 96 of 148 SMs: it is bound by "
                             "L2->SM operand latency and cluster hand-overs, not by the tensor pipe (DESIGN.md section 5); the decoder "
                             "convolutions are the tensor-bound kernels (62 % tensor-pipe active in ncu, profiles/r01_g_conv_full_summary.txt)"},
        "decode": {"ms": dec_s * 1000.0, "tflops": FLOP_DECODE_PER_SAMPLE * B / dec_s / 1e12 if dec_s > 0 else None},
    }
    if world == 1 and not args.no_cpu:
        v, info = cpu_sample_rate(B, denoise_steps=100, decode_rows=32)
        line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": info["threads"], "kind": "port",
                                "sample": "batch %d: 100 of the 1000 reverse steps + decode of 32 samples, extrapolated linearly" % B}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# v4 pixel-space workload (BASELINE configs[4]): secondary line, `--workload v4`
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


def run_pix(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from ldm_b200 import v4      # nothing under oracle/ is imported on this arm

    torch.set_grad_enabled(False)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    sampler = ClockSampler(local_rank)
    sampler.start()
    B, K, W = args.batch, args.steps, args.warmup
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
    clocks = sampler.stop(t_win0, time.time())
    if world > 1:
        t = torch.tensor([t_ms, t_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_ms, t_e2e = float(t[0]), float(t[1])
    if rank != 0:
        return
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
        "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"]},
        "roofline": {"bound": "tensor", "kernel": "sampling loop (one CUDA-graph launch = 1000 steps x %d kernels: 15 tcgen05 implicit-GEMM convolutions, "
                                                  "conv1.0 / out_conv / posterior update)" % (launches // (K * N_STEPS)),
                     "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_sustained"],
                     "traffic": None, "peak_source": pk["src"], "ms_per_launch": t_ms / K,
                     "algorithmic_flop_per_launch": flop},
    }
    if not args.no_cpu:
        v, th = pix_cpu_rate(4, 3)
        line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": th, "kind": "port",
                                "sample": "batch 4, 3 of the 1000 reverse steps, extrapolated linearly (oracle/restate_pix.py, bit-equal to the v4 reference)"}
    print(json.dumps(line))


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="samples per GPU per step")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="v2", choices=["v2", "v3", "v4"],
                    help="v2: BASELINE configs[1]/[2] (default). v3: configs[3], multi-conditional denoiser, --batch rows per GPU (default 128), "
                         "each GPU one reference call. v4: configs[4], pixel-space U-Net at 64x64, --batch images per GPU (default 64)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "v4":
        if "--batch" not in " ".join(sys.argv):
            args.batch = 64
        if args.impl == "reference":
            run_pix_reference_arm(args, rank)
            return
    if args.workload == "v3":
        if "--batch" not in " ".join(sys.argv):
            args.batch = 128
        if args.impl == "reference":
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable": "the v3 workload is a secondary line; the reference arm covers the v2 headline"}))
            return
    elif args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        (run_pix if args.workload == "v4" else run_ours)(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
