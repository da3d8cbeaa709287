#!/usr/bin/env python
"""bench.py — denoising-step throughput of the MToV hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Step      one pass of the hot path over one batch: UNet epsilon-prediction
          (DiffusionWrapper.forward, base.yaml) + the DDIM(eta=1) update of one
          sampling step, for `--chunks-per-gpu` 16-frame 256x256 chunks
          ([B,4,2048] tri-plane latents).  The K timed steps walk the reference's
          50-step schedule (BASELINE.json configs[1]).
value     chunk-steps/s over all ranks, inputs resident in HBM, CUDA-event timed,
          max over ranks; weak scaling (each rank samples its own chunks, one
          all-gather of the final latents inside the timed region when N>1).
e2e       same metric through the public nn.Module API with HOST (pinned) buffers:
          every step copies x/cond/image_cond/t host->device, runs the UNet forward and
          the DDIM update, and reads the updated latent back device->host (sync per step).
roofline  dominant kernel family of one forward, CUDA events around each launch.
cpu_baseline / --impl reference
          the CPU restatement of the reference path (oracle/unet_oracle.py; the
          reference tree itself cannot travel to the GPU box and has no installable
          package) on all host threads, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL's "NCCL version ..." banner under torchrun) are
# sent to stderr, and the result line is written to the saved descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


import torch  # noqa: E402

METRIC = "unet_denoise_steps_per_sec_16f_256px"
UNIT = "chunk-steps/s"
SAMPLING_STEPS = 50


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--chunks-per-gpu", type=int, default=1)
    ap.add_argument("--config", default="base", choices=["base", "longvid", "tiny"])
    ap.add_argument("--cpu-baseline-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def cfg_by_name(name):
    from moditalker_b200 import BASE_UNET_CONFIG, LONGVID_UNET_CONFIG, TINY_UNET_CONFIG
    return {"base": BASE_UNET_CONFIG, "longvid": LONGVID_UNET_CONFIG, "tiny": TINY_UNET_CONFIG}[name]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons, power = [], [], set(), []
        for line in self.f.read().splitlines():
            parts = [s.strip() for s in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(power))
        return out


# ------------------------------------------------------------------------------ CPU arm
def cpu_reference_steps(cfg_name, n_steps, chunks):
    """The reference algorithm on the host cores: oracle forward + DDIM update per step."""
    from moditalker_b200.synth import synth_inputs, synth_noise, synth_state_dict
    from oracle.unet_oracle import Oracle, ddim_time_pairs, ddim_update, schedule
    cores = os.cpu_count() or 1
    cfg = cfg_by_name(cfg_name)
    orc = Oracle(cfg, synth_state_dict(cfg, 0))
    x, cond, ic, _ = synth_inputs(chunks, seed=2)
    sch, pairs = schedule(), ddim_time_pairs(1000, SAMPLING_STEPS)
    img = x.clone()
    noise = synth_noise(img.shape, 3, "bench")
    def step(i):
        nonlocal img
        time_, tn = pairs[i % (len(pairs) - 1)]
        eps = orc.forward(img, cond, ic, torch.full((chunks,), time_, dtype=torch.long)).float()
        img = ddim_update(img, eps, noise, sch, time_, tn)
    # Use the thread count that is FASTEST on this host: on many-core boxes torch's CPU ops on these
    # small tensors get slower past a few dozen threads (measured: 128 threads = 92 s/step, ~100x slower).
    try:
        cores = min(cores, len(os.sched_getaffinity(0)))
    except AttributeError:
        pass
    best_t, best_dt = None, None
    for nt in sorted({min(cores, c) for c in (4, 8, 16, 32, 64)}):
        torch.set_num_threads(nt)
        t0 = time.perf_counter(); step(0); dt = time.perf_counter() - t0     # includes first-touch warm-up
        if best_dt is not None and dt > 3.0 * best_dt:
            break                             # far slower with more threads: stop probing
        t0 = time.perf_counter(); step(0); dt = time.perf_counter() - t0
        if best_dt is None or dt < best_dt:
            best_t, best_dt = nt, dt
        elif dt > 1.3 * best_dt:
            break
    torch.set_num_threads(best_t)
    cores = best_t
    img = x.clone()
    t0 = time.perf_counter()
    for i in range(n_steps):
        step(i + 1)
    dt = time.perf_counter() - t0
    return {"value": chunks * n_steps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n_steps} denoising steps (oracle UNet forward fp32 + DDIM update), {cfg_name}.yaml, B={chunks}, "
                      f"torch CPU {torch.get_num_threads()} threads (fastest of a sweep up to {os.cpu_count()} host cores)", "ms_per_step": 1e3 * dt / n_steps}


def run_reference_arm(args, rank):
    if rank != 0:
        return
    W, K = max(args.warmup, 0), max(args.steps, 1)
    K_eff = min(K, 8)     # bounded sample: each CPU step is ~0.2-1 s
    cb = cpu_reference_steps(args.config, K_eff, args.chunks_per_gpu)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"MToV DDIM denoising step, {args.config}.yaml UNet, {args.chunks_per_gpu} chunk(s) of 16 frames "
                               f"256x256 ([B,4,2048] tri-plane latent)", "timed_steps_cpu": K_eff,
                   "note": "CPU restatement of the reference path (oracle port); rank 0 only"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------ GPU arm
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    import __graft_entry__
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if local_rank == 0:
        __graft_entry__.build()          # no-op when the in-tree .so is current
    if world > 1:
        dist.barrier()
    from moditalker_b200 import DDPM, DiffusionWrapper, UNetModel, _lib
    from moditalker_b200.synth import synth_inputs, synth_state_dict

    cfg = cfg_by_name(args.config)
    B = args.chunks_per_gpu
    model = DiffusionWrapper(UNetModel(**cfg))
    model.load_state_dict(synth_state_dict(cfg, 0, "diffusion_model."), strict=True)
    model = model.to(dev).eval()
    ddpm = DDPM(model, channels=4, image_size=32, sampling_timesteps=SAMPLING_STEPS, w=0.0).to(dev)
    # each rank samples its own chunks (weak scaling): inputs depend on the global chunk index
    x_h, cond_h, ic_h, _ = synth_inputs(B * world, seed=2)
    sl = slice(rank * B, (rank + 1) * B)
    x_h, cond_h, ic_h = x_h[sl].contiguous().pin_memory(), cond_h[sl].contiguous().pin_memory(), ic_h[sl].contiguous().pin_memory()
    cond, ic = cond_h.to(dev), ic_h.to(dev)
    pairs = ddpm.time_pairs()
    lib, h = model.diffusion_model.native_handle(dev)
    stream = torch.cuda.current_stream(dev)
    tconds = {t: torch.full((B,), t, device=dev, dtype=torch.long) for t, _ in pairs}
    img = x_h.to(dev).clone()
    W, K = max(args.warmup, 3), max(args.steps, 1)

    def dev_step(i):
        time_, tn = pairs[i % (len(pairs) - 1)]      # never the final (noise-free) pair: every step is a full update
        eps = model(img, cond, ic, tconds[time_])
        noise = torch.randn_like(img)
        sr, srm1, san, c, sigma = ddpm.step_scalars(time_, tn)
        _lib.check(lib.mtv_ddim_step(h, img.data_ptr(), eps.data_ptr(), noise.data_ptr(), img.numel(), sr, srm1, san, c,
                                     sigma, 0, stream.cuda_stream), "mtv_ddim_step")

    with torch.no_grad():
        for i in range(W):
            dev_step(i)
        torch.cuda.synchronize()
        gathered = torch.empty((world * B, 4, 2048), device=dev) if world > 1 else None
        if world > 1:
            dist.all_gather_into_tensor(gathered, img)     # warm the communicator
            dist.barrier()
        torch.cuda.synchronize()
        clk = ClockSampler(local_rank) if rank == 0 else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(K):
            dev_step(W + i)
        if world > 1:
            dist.all_gather_into_tensor(gathered, img)     # the path's single collective (final latents)
        e1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms_total = float(ms.item())
        clocks = clk.stop() if clk else None

        # ---- e2e: host buffers in, host result out, every step
        eps_h = torch.empty((B, 4, 2048), dtype=torch.float32).pin_memory()
        t_h = torch.full((B,), 500, dtype=torch.long).pin_memory()
        x_d, c_d, ic_d, t_d = torch.empty_like(img), torch.empty_like(cond), torch.empty_like(ic), torch.empty((B,), device=dev, dtype=torch.long)

        def e2e_step(i):
            time_, tn = pairs[i % (len(pairs) - 1)]
            t_h.fill_(time_)
            x_d.copy_(x_h, non_blocking=True); c_d.copy_(cond_h, non_blocking=True)
            ic_d.copy_(ic_h, non_blocking=True); t_d.copy_(t_h, non_blocking=True)
            eps = model(x_d, c_d, ic_d, t_d)
            noise = torch.randn_like(x_d)
            sr, srm1, san, c, sigma = ddpm.step_scalars(time_, tn)
            _lib.check(lib.mtv_ddim_step(h, x_d.data_ptr(), eps.data_ptr(), noise.data_ptr(), x_d.numel(), sr, srm1, san, c,
                                         sigma, 0, stream.cuda_stream), "mtv_ddim_step")
            eps_h.copy_(x_d, non_blocking=True)           # the step's result: the updated latent x_{t-1}
            stream.synchronize()                          # the caller reads it on the host

        for i in range(3):
            e2e_step(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0.record(stream)
        for i in range(K):
            e2e_step(i)
        e1.record(stream)
        torch.cuda.synchronize()
        ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        e2e_ms = float(ms2.item())
        h2d = x_h.numel() * 4 + cond_h.numel() * 4 + ic_h.numel() * 4 + t_h.numel() * 8
        d2h = eps_h.numel() * 4

        # ---- per-kernel-family timing of one forward (events around every launch, same stream)
        fam = {}
        if rank == 0:
            tc = tconds[pairs[0][0]]
            os.environ["MTV_PROFILE_REPS"] = "10"     # each launch repeated back to back: steady-state per-launch time
            for _ in range(3):
                _, rows = model.diffusion_model.profile_forward(img, cond, ic, tc)
            acc = {}
            reps = 3
            for _ in range(reps):
                _, rows = model.diffusion_model.profile_forward(img, cond, ic, tc)
                for name, us, flops, byts in rows:
                    k = name.split(":")[0]
                    a = acc.setdefault(k, [0.0, 0.0, 0.0, 0])
                    a[0] += us; a[1] += flops; a[2] += byts; a[3] += 1
            for k, (us, fl, by, n) in acc.items():
                fam[k] = {"us_per_forward": us / reps, "launch_groups": n // reps, "gflop": fl / reps / 1e9, "mbytes": by / reps / 1e6}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    info = model.diffusion_model.plan_info(B)
    value = world * B * K / (ms_total / 1e3)
    e2e_value = world * B * K / (e2e_ms / 1e3)
    dom = max((k for k in fam if k not in ("copy_t", "copy_out", "pack_in")), key=lambda k: fam[k]["us_per_forward"])
    d = fam[dom]
    if dom in ("conv", "attn", "conv_tc", "attn_tc"):
        ach = d["gflop"] / (d["us_per_forward"] * 1e-6) / 1e3   # TFLOP/s
        # the family is timed launch by launch, each replayed back to back on its own: the BURST peak is the denominator
        roof = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                "frac": ach / pk["bf16_tflops"], "traffic": None,
                "peak_source": f"{pk['source']} (MEASURED_PEAKS.json bf16_tflops, burst: kernels timed alone)",
                "note": "achieved = algorithmic FLOPs (2*M*N*K, one product = 2 FLOP) of all launches of the family in one forward / "
                        "their summed per-launch time (each launch timed as a node of a private CUDA graph, CUDA events on the "
                        "launching stream); the kernel issues 3 bf16 MMAs per product (split-bf16), so tensor-pipe activity is 3x "
                        "this fraction; traffic = dram__bytes_read+write of the family per launch from the committed ncu list "
                        "(profiles/r01_s2_launches_b*.md); what paces the main loop: profiles/r01_s2_mainloop_skip.md"}
    else:
        ach = d["mbytes"] / (d["us_per_forward"] * 1e-6) / 1e3  # GB/s
        roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": ach / pk["hbm_gbs"], "traffic": None, "peak_source": f"{pk['source']} (MEASURED_PEAKS.json hbm_gbs)"}
    roof["us_per_forward"] = d["us_per_forward"]
    roof["launches_per_forward"] = d["launch_groups"]
    # DRAM traffic of the same kernel family from the committed ncu capture (profiles/r01_traffic.json, written by
    # scripts/summarize_traffic.py from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`); per launch like `achieved`
    tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tp):
        tr = json.load(open(tp)).get(f"B{B}", {}).get(dom)
        if tr:
            roof["traffic"] = tr["dram_bytes"] / max(1, tr["launches"])
            roof["traffic_per_forward"] = tr["dram_bytes"]
            roof["algorithmic_bytes_per_forward"] = d["mbytes"] * 1e6
    # whole-step HBM view (0.54 GB algorithmic bytes per step at B=1)
    step_bytes = info["weight_bytes"] + B * (10.6e6 + 0.14e6)
    roof["step_hbm_frac"] = (step_bytes / ((ms_total / K) * 1e-3)) / 1e9 / pk["hbm_gbs"]

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu = cpu_reference_steps(args.config, args.cpu_baseline_steps, B)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": f"MToV 50-step DDIM(eta=1) schedule, {args.config}.yaml UNet, {B} chunk(s)/GPU of 16 frames 256x256 "
                        f"([B,4,2048] tri-plane latent), random cond/image_cond, seeded random weights",
            "arithmetic": "fp32 in/out; contractions as 3 split-bf16 tcgen05 products with fp32 (TMEM) accumulation, softmax / norms in fp32",
            "chunks_per_gpu": B, "global_chunks": B * world, "frames_per_sec_at_50_steps": 16.0 * value / SAMPLING_STEPS,
            "l2": f"per-step working set {step_bytes / 1e9:.2f} GB (weights re-streamed every step) > 126 MB L2; no explicit flush",
            "parallelism": f"chunk-sharded x{world}" + (", one all-gather of final latents in the timed region" if world > 1 else ""),
        },
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / K, "api": "DiffusionWrapper.forward + DDIM update, pinned host buffers in, updated latent out, sync per step"},
        "gpu_launches": int(K * (info["launches"] + 1)),
        "roofline": roof,
        "kernel_families_us": {k: round(v["us_per_forward"], 1) for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["us_per_forward"])},
        "cpu_baseline": cpu,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
