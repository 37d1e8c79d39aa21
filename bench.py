#!/usr/bin/env python3
"""Benchmark of the ZipVoice sampler hot path (BASELINE.json metric: generated mel frames/s of
16-step ZipVoice sampling; p50 RTF).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm: the oracle port of the
                                                              # reference on the host cores

Workload (config C3 of BASELINE.json, per GPU): 64 utterances, 3 s prompt (281 frames, 45 tokens)
+ ~10 s target (938 frames, 150 tokens), ZipVoice 123M with seeded synthetic weights, 16 Euler
steps, classifier-free guidance 1.0, t_shift 0.5.  Weak scaling: every rank samples its own 64
utterances, no collective on the data path (the output mels are gathered at the end).
One "step" = one `sample` call over the rank's batch.  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import collections
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "generated mel frames/s, 16-step ZipVoice sampling (CFG)"
UNIT = "frames/s"
FRAME_SEC = 256.0 / 24000.0
PROMPT_FRAMES, TARGET_FRAMES, PROMPT_TOKENS, TOKENS = 281, 938, 45, 150
NUM_STEP, GUIDANCE, T_SHIFT = 16, 1.0, 0.5


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tflops=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm_gbs=6650.0, tflops=1400.0, src="fallback")   # B200_PROFILING.md fallback


def workload_config(batch_per_gpu, n_gpus, workspace_gb=None):
    return {"workload": f"C3: ZipVoice 123M 16-step CFG sampling, {batch_per_gpu} utterances/GPU "
                        f"({PROMPT_FRAMES}+{TARGET_FRAMES} frames), guidance {GUIDANCE}, t_shift {T_SHIFT}",
            "utterances_per_gpu": batch_per_gpu, "global_utterances": batch_per_gpu * n_gpus,
            "prompt_frames": PROMPT_FRAMES, "target_frames": TARGET_FRAMES, "num_step": NUM_STEP,
            "parallelism": f"utterance-sharded x{n_gpus}, no data-path collective",
            "l2": ("working set (%s workspace per rank) exceeds the 126 MB L2; no explicit flush"
                   % (f"{workspace_gb:.1f} GB" if workspace_gb else "multi-GB"))}


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_steps(n_steps: int, n_warm: int):
    """Times the oracle (CPU restatement of the reference) on a bounded sample: one utterance
    (281+938 frames), one CFG Euler step (two decoder rows) per timed step."""
    import torch
    from oracle import zipvoice_oracle as orc
    from zipvoice_b200.config import ZipVoiceConfig
    from zipvoice_b200.synth import synth_state_dict
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = ZipVoiceConfig("zipvoice")
    model = orc.OracleModel(cfg, synth_state_dict(cfg, 0))
    T = PROMPT_FRAMES + TARGET_FRAMES
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, T, 100, generator=g)
    text = torch.randn(1, T, 100, generator=g) * 0.5
    speech = torch.zeros(1, T, 100)
    speech[:, :PROMPT_FRAMES] = torch.randn(1, PROMPT_FRAMES, 100, generator=g) * 0.3 - 0.5
    mask = torch.zeros(1, T, dtype=torch.bool)
    ts = orc.get_time_steps(0.0, 1.0, NUM_STEP, T_SHIFT)
    times = []
    with torch.inference_mode():
        for i in range(n_warm + n_steps):
            t0 = time.perf_counter()
            v = orc.cfg_velocity(model.sd, model.fc, ts[i % NUM_STEP], x, text, speech, mask, GUIDANCE, False)
            x = x + v * (ts[i % NUM_STEP + 1] - ts[i % NUM_STEP])
            dt = time.perf_counter() - t0
            if i >= n_warm:
                times.append(dt)
    per_step = statistics.mean(times)
    value = TARGET_FRAMES / (NUM_STEP * per_step)
    return value, per_step, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    value, per_step, cores = cpu_reference_steps(max(1, args.steps), max(1, min(args.warmup, 2)))
    sample = (f"1 utterance ({PROMPT_FRAMES}+{TARGET_FRAMES} frames); each timed step = 1 of the 16 CFG Euler "
              f"steps (2 decoder rows), frames/s scaled to 16 steps")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3 * NUM_STEP,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.batch_per_gpu, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------- GPU arm
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons, power = [], [], set(), []
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def run_b200(args):
    import torch
    import torch.distributed as dist
    from zipvoice_b200 import _lib
    from zipvoice_b200.config import ZipVoiceConfig
    from zipvoice_b200.model import build_model
    from zipvoice_b200.synth import synth_state_dict, synth_utterances

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    B = args.batch_per_gpu
    cfg = ZipVoiceConfig("zipvoice")
    model = build_model(cfg, synth_state_dict(cfg, 0), dev, use_cuda_graph=True)
    u = synth_utterances(cfg, batch=B, prompt_frames=PROMPT_FRAMES, target_frames=TARGET_FRAMES,
                         prompt_tokens=PROMPT_TOKENS, tokens=TOKENS, seed=666 + rank)
    kw = dict(num_step=NUM_STEP, guidance_scale=GUIDANCE, t_shift=T_SHIFT)
    pf_host = u["prompt_features"].pin_memory()
    pfl_host = u["prompt_features_lens"].pin_memory()
    tgt_host = u["target_lens"].pin_memory()
    frames_per_step = int(u["target_lens"].sum())
    audio_sec = frames_per_step * FRAME_SEC

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident leg: conditions already in HBM, K calls of solver.sample
    with torch.inference_mode():
        tc, pm = model.forward_text_inference_gt_duration(
            tokens=u["tokens"], features_lens=u["target_lens"].to(dev), prompt_tokens=u["prompt_tokens"],
            prompt_features_lens=u["prompt_features_lens"].to(dev))
        T = tc.shape[1]
        sc = torch.zeros(B, T, cfg.feat_dim, device=dev)
        sc[:, :PROMPT_FRAMES] = u["prompt_features"].to(dev)
        x0 = u["x0"].to(dev)
        lc0 = _lib.launch_count()
        model.solver.sample(x=x0, text_condition=tc, speech_condition=sc, padding_mask=pm, **kw)   # captures the graph
        launches_per_sample = _lib.launch_count() - lc0
        for _ in range(max(args.warmup, 3) - 1):
            model.solver.sample(x=x0, text_condition=tc, speech_condition=sc, padding_mask=pm, **kw)
        barrier()
        clocks = ClockSampler(local)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        evs[0].record()
        for i in range(args.steps):
            x1 = model.solver.sample(x=x0, text_condition=tc, speech_condition=sc, padding_mask=pm, **kw)
            evs[i + 1].record()
        barrier()
        step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
        total_ms = max_over_ranks(evs[0].elapsed_time(evs[-1]))
        clock_info = clocks.stop()
        finite = bool(torch.isfinite(x1).all())

        # ---- end-to-end leg: the public API with HOST inputs (tokens, pinned prompt mel), H2D and D2H inside
        out_host = torch.empty(B, TARGET_FRAMES, cfg.feat_dim).pin_memory()
        lens_host = torch.empty(B, dtype=torch.int64).pin_memory()

        def e2e_step():
            pf = pf_host.to(dev, non_blocking=True)
            pfl = pfl_host.to(dev, non_blocking=True)
            tl = tgt_host.to(dev, non_blocking=True)
            mel, lens, _, _ = model.sample(u["tokens"], u["prompt_tokens"], pf, pfl, features_lens=tl,
                                           duration="real", **kw)
            if world > 1:     # the only collective: gather the output mels (NCCL over NVLink)
                gathered = torch.empty(world, *mel.shape, device=dev, dtype=mel.dtype)
                dist.all_gather_into_tensor(gathered, mel.contiguous())
            out_host.copy_(mel, non_blocking=True)
            lens_host.copy_(lens, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
    h2d = pf_host.numel() * 4 + pfl_host.numel() * 8 + tgt_host.numel() * 8
    d2h = out_host.numel() * 4 + lens_host.numel() * 8

    # ---- roofline: CUDA events around every kernel of one decoder forward (on the launching stream), each
    # kernel against ITS OWN bound: max(algorithmic FLOPs / tensor peak, algorithmic HBM bytes / copy peak).
    # The K <= 512 residual-stream GEMMs and the 12-column SelfAttention P.V products are HBM bound, the
    # feed-forward / gated / NonlinAttention GEMMs tensor bound.
    plan = model.solver.decoders[cfg.feat_dim].plans.get(2 * B, T)
    plan.profile()
    peaks = load_peaks()
    agg = collections.OrderedDict()
    reps = 3
    for _ in range(reps):
        for cat, ms, work, nbytes in plan.profile(with_bytes=True):
            tensor = cat.startswith("gemm") or cat == "attn_weights"
            flops = work if tensor else 0.0
            t_t = flops / (peaks["tflops"] * 1e12) * 1e3
            t_h = nbytes / (peaks["hbm_gbs"] * 1e9) * 1e3
            d = agg.setdefault(cat, dict(n=0, ms=0.0, flops=0.0, bytes=0.0, roof=0.0, roof_t=0.0, roof_h=0.0))
            d["n"] += 1; d["ms"] += ms; d["flops"] += flops; d["bytes"] += nbytes
            d["roof"] += max(t_t, t_h)
            if t_t >= t_h:
                d["roof_t"] += t_t
            else:
                d["roof_h"] += t_h
    tot_ms = sum(d["ms"] for d in agg.values())
    kernels = []
    for cat, d in agg.items():
        bound = "tensor" if d["roof_t"] >= d["roof_h"] else "hbm"
        rate = (d["flops"] / 1e12 if bound == "tensor" else d["bytes"] / 1e9) / (d["ms"] * 1e-3) if d["ms"] > 0 else 0.0
        kernels.append({"kernel": cat, "launches_per_forward": d["n"] // reps, "share_of_forward": d["ms"] / tot_ms,
                        "bound": bound, "achieved": rate, "unit": "TFLOP/s" if bound == "tensor" else "GB/s",
                        "peak": peaks["tflops"] if bound == "tensor" else peaks["hbm_gbs"],
                        # time the kernel's own roofline allows / time measured (mixed bound, launch by launch)
                        "frac": d["roof"] / d["ms"] if d["ms"] > 0 else 0.0,
                        "tflops": d["flops"] / 1e12 / (d["ms"] * 1e-3) if d["ms"] > 0 else 0.0,
                        "gbs": d["bytes"] / 1e9 / (d["ms"] * 1e-3) if d["ms"] > 0 else 0.0})
    # dominant kernel = gemm_kernel (every epilogue variant: linear, gated, P.V share one kernel template)
    gemm = [d for cat, d in agg.items() if cat.startswith("gemm")]
    g_n = sum(d["n"] for d in gemm); g_ms = sum(d["ms"] for d in gemm); g_work = sum(d["flops"] for d in gemm)
    g_roof = sum(d["roof"] for d in gemm); g_bytes = sum(d["bytes"] for d in gemm)
    achieved = g_work / (g_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_r1.json")
    if os.path.exists(tpath):          # avg dram bytes per gemm_kernel launch, from the committed ncu capture
        traffic = json.load(open(tpath)).get("gemm_kernel_dram_bytes_per_launch")
    roofline = {"kernel": "gemm_kernel (tcgen05, all epilogue variants)", "bound": "tensor", "achieved": achieved,
                "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"], "traffic": traffic,
                "algorithmic_flops_per_launch": g_work / g_n, "algorithmic_bytes_per_launch": g_bytes / g_n,
                "avg_launch_ms": g_ms / g_n, "launches_per_forward": g_n // reps,
                # launch by launch against max(tensor time, HBM time): the HBM-bound launches (residual-stream
                # GEMMs with K <= 512, SelfAttention P.V) cannot reach the tensor peak by construction
                "frac_of_own_roofline": g_roof / g_ms,
                "peak_source": f"{peaks['src']} (sustained figure: kernel timed inside a long step)",
                "share_of_forward": g_ms / tot_ms, "forward_ms": tot_ms / reps,
                "forward_frac_of_roofline": sum(d["roof"] for d in agg.values()) / tot_ms}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        v, per_step, cores = cpu_reference_steps(2, 1)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"1 utterance ({PROMPT_FRAMES}+{TARGET_FRAMES} frames), 2 of the 16 CFG Euler steps "
                                  f"timed after 1 warm-up ({per_step:.2f} s/step), scaled to 16 steps"}
    value = frames_per_step * world * args.steps / (total_ms * 1e-3)
    e2e_value = frames_per_step * world * args.steps / e2e_s
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
            "config": workload_config(B, world, plan.workspace_bytes / 1e9),
            "rtf_p50": statistics.median(step_ms) * 1e-3 / audio_sec,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s * 1e3 / args.steps, "rtf": e2e_s / args.steps / audio_sec},
            "gpu_launches": launches_per_sample * args.steps,
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu_baseline, "clocks": clock_info,
            "finite": finite}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch-per-gpu", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
