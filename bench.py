#!/usr/bin/env python
"""bench.py -- audio-seconds/second of the CoVoMix inference hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c3|c2]

One "step" = one pass of the hot path over one batch of synthetic input:
    workload c3 (default; BASELINE.json configs[2], the config the metric is quoted on):
        VoMix 2-stream acoustic model, 64 Euler flow-matching steps (64 NFE x 2 CFG passes),
        B = 8 dialogues of 30 s (N = 1650 frames = 150 prompt + 1500 generated), cond_scale 0.7,
        followed by the HiFi-GAN generator on the 8 x [80, 1500] generated mels -> 8 x 240032 samples.
    workload c2: VoSingle, 32 Euler steps, 10 s monologue (N = 650), B = 1, + vocoder.
    workload c4: the full pipeline per GPU (BASELINE.json configs[3]): CoMix text-to-semantic -> VoMix -> vocoder.
    workload c4p: c4 software-pipelined over batches: T2S of the next batch on its own SM budget next to flow + vocoder.
    workload c5: HiFi-GAN sweep (one JSON line per point).
value = B * 30 s * n_gpus / seconds-per-step (whole job, inputs resident in HBM).
e2e   = the same through the public Python API with pinned HOST inputs and a device->host read of the waveform.

--impl reference times the reference algorithm's CPU implementation (the fp32 PyTorch oracle port,
oracle/covomix_oracle.py; the reference itself is Python and /root/reference does not travel) on a bounded sample.
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
sys.path.insert(0, ROOT)

import torch  # noqa: E402

FRAME_RATE = 50.0
WORKLOADS = {
    "c3": dict(model="vomix", B=8, N=1650, prompt=150, method="euler", n_steps=64,
               name="C3: VoMix 2-stream, 64 Euler steps (128 network passes), 30 s dialogue, batch 8, + HiFi-GAN"),
    "c2": dict(model="vosingle", B=1, N=650, prompt=150, method="euler", n_steps=32,
               name="C2: VoSingle, 32 Euler steps, 10 s monologue, batch 1, + HiFi-GAN"),
    # C3's shape with the REFERENCE DEFAULT solver (acoustic.py:586-591: midpoint, step 1/16 = 32 NFE; SURVEY 8d "step-count naming")
    "c3m": dict(model="vomix", B=8, N=1650, prompt=150, method="midpoint", n_steps=16,
                name="C3 shape, reference-default solver: VoMix 2-stream, midpoint 16 steps (32 NFE, 64 network passes), 30 s dialogue, "
                     "batch 8, + HiFi-GAN"),
    # BASELINE.json configs[3] per GPU: the full pipeline.  CoMix text-to-semantic (200 text tokens -> 1500 positions x 2
    # streams, EOS ignored: random-init weights would stop at a random position) -> VoMix (as C3) -> HiFi-GAN.
    "c4": dict(model="vomix", B=8, N=1650, prompt=150, method="euler", n_steps=64, t2s=dict(S=200, steps=1500),
               name="C4 (per GPU): CoMix T2S (1500 AR steps, 2 streams) -> VoMix 64 Euler steps -> HiFi-GAN, 30 s dialogues, batch 8"),
    # same work, software-pipelined over batches on two streams with an SM budget per stage: the latency-bound T2S loop of
    # batch i+1 (t2s_sms SMs) runs next to the flow sampler + vocoder of batch i (the remaining SMs)
    "c4p": dict(model="vomix", B=8, N=1650, prompt=150, method="euler", n_steps=64, t2s=dict(S=200, steps=1500), t2s_sms=32,
                name="C4 pipelined (per GPU): CoMix T2S of batch i+1 on 32 SMs || VoMix 64 Euler steps + HiFi-GAN of batch i "
                     "on 116 SMs, 30 s dialogues, batch 8"),
}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(tflops=p.get("bf16_tflops_sustained", 1409.8), hbm=p.get("hbm_gbs", 6548.8), src="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(tflops=1400.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for nm, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        # median over the samples taken under load (upper half; the first samples can predate the launch)
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic_per_launch():
    """Mean DRAM read+write bytes per GEMM launch from the committed ncu --set full capture of this workload
    (profiles/r0N_gemm_c3_ncu.md, newest round first; written by tools/summarize_ncu.py); (None, None) if there is no capture."""
    for name in ("r02_gemm_c3_ncu.md", "r01_gemm_c3_ncu.md"):
        path = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(path):
            continue
        vals = []
        for ln in open(path):
            if ln.startswith("- DRAM traffic"):
                try:
                    vals.append(float(ln.split("=")[-1].split()[0]) * 1e6)
                except ValueError:
                    pass
        if vals:
            return sum(vals) / len(vals), name
    return None, None


def config_for(wl, world, t2s_sms=0, flow_sms=None):
    """The `config` object of the JSON line: identical for the B200 arm and the reference arm (the driver compares them)."""
    cfg = {"workload": wl["name"], "global_batch": wl["B"] * world, "seq_len": wl["N"],
           "parallelism": f"utterance-sharded x{world}", "cond_scale": 0.7,
           "sm_budget": {"t2s": t2s_sms, "flow+vocoder": "all other SMs"} if t2s_sms else None,
           "l2": "working set (0.8 GB weights + 1.2 GB activations) larger than L2; no flush needed"}
    if "t2s" in wl:
        cfg["t2s_assumption"] = ("text-to-semantic decodes the 8 dialogues of a batch in ONE call with EOS ignored (random-init weights "
                                 "would stop at a random position); covomix_b200.pipeline.covomix_dialogues decodes one dialogue per "
                                 "call like the reference, because its EOS rule couples the rows of a batch")
    return cfg


def make_inputs(wl, device=None, seed=30):
    from covomix_b200 import synthetic as syn
    cfg = syn.VOMIX if wl["model"] == "vomix" else syn.VOSINGLE
    ids, cond, y0, mask = syn.synthetic_flow_inputs(cfg, wl["B"], wl["N"], prompt=wl["prompt"], seed=seed)
    return cfg, ids, cond, y0, mask


def algorithmic_flops(cfg, wl):
    """per token per pass 2*(L_lin + 16384*N); vocoder 281.3 MFLOP per mel frame."""
    N, B = wl["N"], wl["B"]
    inner = cfg.heads * cfg.dim_head
    lin = cfg.embed_in * cfg.dim + 31 * cfg.dim + cfg.depth * (cfg.dim * 3 * inner + inner * cfg.dim + 2 * cfg.dim * cfg.dim * cfg.ff_mult) \
        + (cfg.depth // 2) * 2 * cfg.dim * cfg.dim + cfg.dim * cfg.dim_x
    attn = cfg.depth * 2 * N * inner
    per_pass = 2.0 * (lin + attn) * N * B
    nfe = wl["n_steps"] * (2 if wl["method"] == "midpoint" else 1)
    flow = per_pass * nfe * 2
    voc = 281.3e6 * (N - wl["prompt"]) * B
    return flow, voc


# ======================================================================================== reference arm (CPU)
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import covomix_b200  # noqa: F401
    from covomix_b200 import synthetic as syn
    from oracle import covomix_oracle as orc
    wl = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = syn.VOMIX if wl["model"] == "vomix" else syn.VOSINGLE
    sd = syn.synthetic_flow_state_dict(cfg, 1234)
    hsd = syn.synthetic_hifigan_state_dict(syn.HIFIGAN_COVOMIX, 1234)
    # bounded sample: ONE item of the batch, ONE CFG velocity evaluation (2 network passes) + the vocoder on one item;
    # the full step is B items x n_eval evaluations (+ B vocoder passes) of exactly this work, so seconds-per-step is
    # extrapolated linearly (the reference processes items one at a time: dialogue_generation.py:283).
    wl1 = dict(wl, B=1)
    _, ids, cond, y0, _ = make_inputs(wl1)
    gen_frames = wl["N"] - wl["prompt"]
    mel = syn.synthetic_logmel(torch.Generator().manual_seed(31), 1, 80, gen_frames)
    n_eval = wl["n_steps"] * (2 if wl["method"] == "midpoint" else 1)

    def one():
        t0 = time.perf_counter()
        with torch.inference_mode():
            orc.velocity_cfg(sd, cfg, y0, ids, cond, torch.tensor(0.5), 0.7)
        t1 = time.perf_counter()
        orc.hifigan_forward(hsd, syn.HIFIGAN_COVOMIX, mel)
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1

    for _ in range(max(1, min(args.warmup, 1))):
        one()
    steps = max(1, min(args.steps, 3))
    tv, th = 0.0, 0.0
    for _ in range(steps):
        a, b = one()
        tv += a
        th += b
    tv /= steps
    th /= steps
    sec_per_step = wl["B"] * (n_eval * tv + th)
    if "t2s" in wl:
        sec_per_step += wl["B"] * wl["t2s"]["steps"] * cpu_t2s_step_seconds(wl)
    audio_s = wl["B"] * gen_frames / FRAME_RATE
    value = audio_s / sec_per_step
    sample = (f"1 of {wl['B']} items, 1 of {n_eval} CFG velocity evaluations ({tv:.2f} s) + vocoder on 1 item ({th:.2f} s), "
              f"{steps} reps; step time extrapolated = B*(n_eval*t_eval + t_voc)")
    if "t2s" in wl:
        sample += f" + B*{wl['t2s']['steps']} text-to-semantic decoding steps timed on a 48-step sample (oracle, B = 1)"
    line = {
        "impl": "reference", "metric": "audio-seconds/sec (RTF)", "value": value, "unit": "audio-s/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_for(wl, max(1, args.gpus), int(os.environ.get("COVO_T2S_SMS", wl.get("t2s_sms", 0))),
                             None),
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "ode_step_ms": tv * 1e3 * wl["B"],
    }
    if args.workload == "c3":
        line["cpu_baseline"]["c2_measured"] = cpu_c2_measured()
    print(json.dumps(line), flush=True)


# ======================================================================================== B200 arm
def run_b200(args):
    """Default run: the headline workload (one JSON line, the bench contract) -- and, for the default workload c3, the other
    BASELINE configs measured in the same process right after it and attached under `other_configs` (C2, C4 pipelined,
    the C5 vocoder sweep), so that every BASELINE config is on the driver's record at every N."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    ctx = dict(world=world, rank=rank, local=local, dev=dev, dist=dist)
    line = bench_workload(args, args.workload, ctx, detail=True)
    if args.workload == "c3" and not args.no_other_configs:
        others = {}
        for key, steps in (("c2", 5), ("c3m", 2), ("c4p", 2)):
            try:
                sub = bench_workload(argparse.Namespace(**{**vars(args), "steps": steps, "no_cpu_baseline": True}), key, ctx,
                                     detail=False)
                if sub is not None:
                    others[key] = {k: sub[k] for k in ("value", "unit", "ms_per_step", "n_gpus", "steps", "warmup", "config", "e2e",
                                                       "ode_step_ms", "clocks", "gpu_launches") if k in sub}
            except Exception as e:                      # the headline line must survive a failure of an attached config
                others[key] = {"error": f"{type(e).__name__}: {e}"}
        try:
            sweep = run_vocoder_sweep(args, ctx, quiet=True)
            if sweep is not None:
                others["c5"] = sweep
        except Exception as e:
            others["c5"] = {"error": f"{type(e).__name__}: {e}"}
        if line is not None:
            line["other_configs"] = others
    if line is not None:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return line


def bench_workload(args, wl_key, ctx, detail=True):
    import covomix_b200  # noqa: F401
    from covomix_b200 import _native as nat, synthetic as syn
    from covomix_b200.flow import B200FlowSampler
    from covomix_b200.vocoder import B200Generator

    world, rank, local, dev, dist = ctx["world"], ctx["rank"], ctx["local"], ctx["dev"], ctx["dist"]
    wl = WORKLOADS[wl_key]
    cfg, ids_h, cond_h, y0_h, mask_h = make_inputs(wl, seed=30 + rank)    # each rank = its own shard of utterances
    # weights: identical on every rank (seeded); rank 0 could equally broadcast them (NCCL) -- see DESIGN.md
    n_sms = torch.cuda.get_device_properties(dev).multi_processor_count
    t2s_sms = int(os.environ.get("COVO_T2S_SMS", wl.get("t2s_sms", 0)))
    flow_sms = n_sms - t2s_sms if t2s_sms else None
    sampler = B200FlowSampler(syn.synthetic_flow_state_dict(cfg, 1234), cfg, dev, torchdiffeq_ode_method=wl["method"],
                              ode_step_size=1.0 / wl["n_steps"], sm_limit=flow_sms)
    gen = B200Generator(syn.synthetic_hifigan_state_dict(syn.HIFIGAN_COVOMIX, 1234), syn.HIFIGAN_COVOMIX, dev, sm_limit=flow_sms)
    B, N, prompt = wl["B"], wl["N"], wl["prompt"]
    gen_frames = N - prompt
    audio_s_per_step = B * gen_frames / FRAME_RATE

    ids_d, cond_d, y0_d = ids_h.to(dev), cond_h.to(dev), y0_h.to(dev)
    ids_p, cond_p = ids_h.pin_memory(), cond_h.pin_memory()

    t2s = text_d = text_p = None
    if "t2s" in wl:
        from covomix_b200.t2s import B200TextToSemantic
        t2s = B200TextToSemantic(syn.synthetic_t2s_state_dict(syn.COMIX, 1234), syn.COMIX, dev, sm_limit=t2s_sms or None)
        text_h = syn.synthetic_text_ids(syn.COMIX, B, wl["t2s"]["S"], seed=40 + rank, ragged=True)
        text_d, text_p = text_h.to(dev), text_h.pin_memory()
        assert wl["t2s"]["steps"] == gen_frames

    def semantic_ids(text, prompt_ids):
        """comix_pred + the id assembly of dialogue_generation.py:306-313: the two generated streams follow the prompt's."""
        L = wl["t2s"]["steps"]
        tgt = t2s.generate(text, max_length=L, ignore_eos=True)                  # [B, 2L]: stream 1 | stream 2
        new = torch.stack((tgt[:, :L], tgt[:, L:]), dim=-1).clamp_(min=0, max=501)
        return torch.cat((prompt_ids[:, :prompt], new), dim=1)

    def step_device():
        if pipelined:
            return step_pipe(text_d, ids_d, cond_d, y0_d)
        ids = semantic_ids(text_d, ids_d) if t2s is not None else ids_d
        mel = sampler.sample(phoneme_ids=ids, cond=cond_d, cond_scale=0.7, y0=y0_d)
        voc_in = mel[:, prompt:, :].permute(0, 2, 1)           # what the scripts do: sampled[:, mask].permute(0,2,1)
        return gen(voc_in)

    wav_host = torch.empty(B, 1, gen.out_len(gen_frames), dtype=torch.float32).pin_memory()

    pipelined = bool(t2s_sms) and t2s is not None
    if pipelined:
        # Software pipeline over batches: the host first enqueues flow + vocoder of the CURRENT batch on stream B (CUDA-graph
        # launches return at once), then runs the text-to-semantic loop of the NEXT batch on stream A (its one host sync is
        # where the host waits), then joins B.  One step still consumes one batch of text and emits one batch of audio.
        s_a, s_b = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        state = {"ids": semantic_ids(text_d, ids_d)}
        torch.cuda.synchronize()

        def step_pipe(text, prompt_ids, cond, y0):
            cur = torch.cuda.current_stream()
            s_a.wait_stream(cur)
            s_b.wait_stream(cur)
            with torch.cuda.stream(s_b):
                mel = sampler.sample(phoneme_ids=state["ids"], cond=cond, cond_scale=0.7, y0=y0)
                wav = gen(mel[:, prompt:, :].permute(0, 2, 1))
            with torch.cuda.stream(s_a):
                nxt = semantic_ids(text, prompt_ids)
            cur.wait_stream(s_a)
            cur.wait_stream(s_b)
            state["ids"] = nxt
            return wav

    def step_e2e():
        i = ids_p.to(dev, non_blocking=True)
        c = cond_p.to(dev, non_blocking=True)
        if pipelined:
            wav = step_pipe(text_p.to(dev, non_blocking=True), i, c, None)
            wav_host.copy_(wav, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return wav_host
        if t2s is not None:
            i = semantic_ids(text_p.to(dev, non_blocking=True), i)
        mel = sampler.sample(phoneme_ids=i, cond=c, cond_scale=0.7)   # y0 = torch.randn_like on device, as the reference
        wav = gen(mel[:, prompt:, :].permute(0, 2, 1))
        wav_host.copy_(wav, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return wav_host

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, sampler_clock=None):
        for _ in range(warmup):
            fn()
        barrier()
        if sampler_clock:
            sampler_clock.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler_clock.stop() if sampler_clock else None
        if dist is not None:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, clocks

    W, K = max(args.warmup, 3), args.steps
    ms_total, clocks = timed(step_device, K, W, ClockSampler(local))
    ms_step = ms_total / K
    value = audio_s_per_step * world / (ms_step / 1e3)

    ms_e2e_total, _ = timed(step_e2e, K, 1)
    ms_e2e = ms_e2e_total / K
    e2e_value = audio_s_per_step * world / (ms_e2e / 1e3)
    h2d = ids_p.numel() * 8 + cond_p.numel() * 4 + (text_p.numel() * 8 if text_p is not None else 0)
    d2h = wav_host.numel() * 4

    line = None
    if rank == 0:
        # ---- ODE-step latency (one CFG velocity evaluation = 2 network passes batched as 2B), CUDA events
        x = y0_d
        for _ in range(2):
            sampler.velocity(x, times=0.5, phoneme_ids=ids_d, cond=cond_d, cond_scale=0.7)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            sampler.velocity(x, times=0.5, phoneme_ids=ids_d, cond=cond_d, cond_scale=0.7)
        e1.record()
        torch.cuda.synchronize()
        ode_step_ms = e0.elapsed_time(e1) / 3
        launches = (sampler.last_launches() + gen.launches_per_forward()) * K + (t2s.launches_per_generate() * K if t2s is not None else 0)
        line = {
            "metric": "audio-seconds/sec (RTF)", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 operands / f32 accumulate (flow), fp16 operands / f32 accumulate (vocoder)", "data": "synthetic",
            "config": config_for(wl, world, t2s_sms, flow_sms),
            "clocks": clocks, "gpu_launches": launches * world,
            "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e},
            "ode_step_ms": ode_step_ms,
        }

    if rank == 0 and detail:
        # ---- roofline of the dominant kernel: one extra instrumented step (graphs bypassed, every launch of the
        # step bracketed by CUDA events on the launching stream), same workload, same process
        with nat.profile() as prof:
            step_device()
        torch.cuda.synchronize()
        pr = prof.result
        total_kernel_ms = sum(v[0] for v in pr.values())
        g_ms, g_flops, g_n = pr["gemm_tc"]
        v_ms, v_flops, v_n = pr["gemm_tc_vocoder"]
        pk = peaks()
        achieved = g_flops / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
        flow_fl, voc_fl = algorithmic_flops(cfg, wl)
        roofline = {
            "kernel": "gemm_tc_kernel (tcgen05 implicit-GEMM), launches of the velocity net's Linear layers",
            "bound": "tensor", "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s",
            "frac": achieved / pk["tflops"], "peak_source": pk["src"],
            "traffic": ncu_traffic_per_launch()[0] if wl_key == "c3" else None,
            "traffic_note": f"mean dram__bytes_read+write per GEMM launch over the launches of profiles/{ncu_traffic_per_launch()[1]} (bytes)",
            "launches": g_n, "avg_launch_ms": g_ms / max(g_n, 1), "algorithmic_flops_per_launch": g_flops / max(g_n, 1),
            "share_of_step_kernel_time": g_ms / total_kernel_ms if total_kernel_ms else None,
            "how": "one extra instrumented step after the timed region: CUDA events around every launch on the launching stream",
            "classes_ms": {k: round(v[0], 3) for k, v in pr.items()},
            "attention": {"achieved_tflops": pr["attention_tc"][1] / (pr["attention_tc"][0] * 1e-3) / 1e12 if pr["attention_tc"][0] else None,
                          "launches": pr["attention_tc"][2]},
            "vocoder_gemm": {"achieved_tflops": v_flops / (v_ms * 1e-3) / 1e12 if v_ms else None, "launches": v_n, "ms": v_ms},
            "whole_step": {"algorithmic_tflop": (flow_fl + voc_fl) / 1e12,
                           "achieved_tflops": (flow_fl + voc_fl) / (ms_step * 1e-3) / 1e12,
                           "frac_of_peak": (flow_fl + voc_fl) / (ms_step * 1e-3) / 1e12 / pk["tflops"]},
        }
        t2s_info = None
        if t2s is not None:
            d_ms = pr["t2s_decode"][0]
            steps_t = wl["t2s"]["steps"]
            gbs = t2s.weight_bytes_per_step() * steps_t / (d_ms * 1e-3) / 1e9 if d_ms else None
            t2s_info = {"kernel": "t2s_decode_kernel (persistent cooperative kernel, whole AR loop in one launch)",
                        "ms": d_ms, "us_per_step": d_ms * 1e3 / steps_t, "tokens_per_s": B * 2 * steps_t / (d_ms * 1e-3) if d_ms else None,
                        "bound": "latency (34 grid barriers per step) over weight streaming",
                        "weight_bytes_per_step": t2s.weight_bytes_per_step(), "weight_stream_gbs": gbs,
                        "frac_of_hbm_peak": gbs / pk["hbm"] if gbs else None, "share_of_step_kernel_time": d_ms / total_kernel_ms}

        # ---- CPU baseline on this box's host cores (oracle port, bounded sample)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline(wl, cfg)

        line["roofline"] = roofline
        if t2s_info is not None:
            line["t2s"] = t2s_info
        if cpu is not None:
            line["cpu_baseline"] = cpu
    # free this workload's device memory before the next one is measured in the same process
    sampler.close()
    gen.close()
    if t2s is not None:
        t2s.close()
    del sampler, gen, t2s
    torch.cuda.empty_cache()
    return line


def vocoder_traffic_model(B, T, fused_last):
    """Bytes moved by the layer-by-layer schedule (what the implementation does): every conv reads its 16-bit input and
    writes a 16-bit and/or fp32 output (+ fp32 residual read); the fused last stage reads its input once (x halo)."""
    c_pad, t_len, byt = [512, 256, 128, 64, 64], T, 0.0
    byt += B * T * (80 * 4 + 128 * 2 + 128 * 2 + c_pad[0] * 2)
    for i, (u, k) in enumerate(zip((5, 4, 4, 2), (8, 8, 4, 4))):
        t_out = (t_len - 1) * u - 2 * ((k - u) // 2) + k
        n = B * t_out * c_pad[i + 1]
        byt += B * t_len * c_pad[i] * 2 + n * 6                      # upsample: read in, write f32 + 16-bit
        if fused_last and i == 3:
            byt += B * t_out * (32 * 4 * 382 / 256 + 4)             # stage input once (x halo) + waveform
        else:
            byt += 3 * (3 * (n * 2 + n * 2) + 3 * (n * 2 + n * 4 + n * 4 + n * 2))   # 3 resblocks x 3 x (c1: r+w, c2: r + res + w f32 + w h)
            byt += 3 * n * 4 + n * 2                                  # stage mean
        t_len = t_out
    if not fused_last:
        byt += B * t_len * (64 * 2 * 1 + 4)
    return byt


def run_vocoder_sweep(args, ctx=None, quiet=False):
    """BASELINE config C5: HiFi-GAN throughput sweep, mel length T in {256, 1024, 4096, 16384} x batch B in {1, 2, 4, 8, 16, 32}
    (points with B*T > 131072 frames are skipped: they are several copies of a smaller point), on 1 GPU or -- under
    torchrun -- on every rank at once (weak scaling: each rank runs the same point on its own mels; value = sum over ranks /
    max-over-ranks time).  One JSON line per point on rank 0 unless `quiet`.  FLOPs are algorithmic (281.3 MFLOP per mel
    frame); `model_gbs` is a hand traffic model of the schedule, NOT a measurement -- the measured DRAM bytes of the bench
    shape are in profiles/ (ncu dram__bytes)."""
    import covomix_b200  # noqa: F401
    from covomix_b200 import _native as nat, synthetic as syn
    from covomix_b200.vocoder import B200Generator
    own_ctx = ctx is None
    if own_ctx:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        rank = int(os.environ.get("RANK", "0"))
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dev = torch.device("cuda", local)
        dist = None
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=dev)
        ctx = dict(world=world, rank=rank, local=local, dev=dev, dist=dist)
    world, rank, dev, dist = ctx["world"], ctx["rank"], ctx["dev"], ctx["dist"]
    gen = B200Generator(syn.synthetic_hifigan_state_dict(syn.HIFIGAN_COVOMIX, 1234), syn.HIFIGAN_COVOMIX, dev)
    pk = peaks()
    g = torch.Generator().manual_seed(5 + rank)
    fused_last = not os.environ.get("COVO_HIFIGAN_NO_FUSED")        # last stage = one tile-resident kernel
    out = []
    for T in (256, 1024, 4096, 16384):
        for B in (1, 2, 4, 8, 16, 32):
            if B * T > 32 * 4096:
                continue
            mel = syn.synthetic_logmel(g, B, 80, T).to(dev)
            for _ in range(3):
                gen(mel)
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = max(3, args.steps)
            e0.record()
            for _ in range(reps):
                gen(mel)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            if dist is not None:
                t = torch.tensor([ms], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            kernel_ms = None
            if not quiet:
                with nat.profile() as prof:
                    gen(mel)
                kernel_ms = {k: round(v[0], 3) for k, v in prof.result.items() if v[2]}
            flops = 281.3e6 * T * B * world
            byt = vocoder_traffic_model(B, T, fused_last) * world
            line = {"metric": "audio-seconds/sec (RTF), HiFi-GAN only", "workload": "C5 vocoder sweep", "T": T, "B": B, "n_gpus": world,
                    "ms": ms, "value": world * B * T / FRAME_RATE / (ms * 1e-3), "unit": "audio-s/s",
                    "achieved_tflops": flops / (ms * 1e-3) / 1e12,
                    "frac_of_tensor_peak": flops / (ms * 1e-3) / 1e12 / (pk["tflops"] * world),
                    "model_gbs": byt / (ms * 1e-3) / 1e9, "model_frac_of_hbm_peak": byt / (ms * 1e-3) / 1e9 / (pk["hbm"] * world),
                    "launches": gen.launches_per_forward()}
            if kernel_ms is not None:
                line["kernel_ms"] = kernel_ms
            if quiet:
                line = {k: (round(v, 4) if isinstance(v, float) else v) for k, v in line.items()
                        if k in ("T", "B", "n_gpus", "ms", "value", "achieved_tflops", "frac_of_tensor_peak")}
            out.append(line)
            if rank == 0 and not quiet:
                print(json.dumps(line), flush=True)
    gen.close()
    del gen
    torch.cuda.empty_cache()
    if own_ctx and dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return out if rank == 0 else None


def cpu_c2_measured():
    """One MEASURED (not extrapolated) end-to-end CPU number: the full BASELINE configs[1] job -- VoSingle, 32 Euler steps
    (64 network passes), N = 650, B = 1, then the vocoder on the 500 generated frames -- through the oracle port on all
    host cores.  About 30-40 s."""
    from covomix_b200 import synthetic as syn
    from oracle import covomix_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wl = WORKLOADS["c2"]
    cfg = syn.VOSINGLE
    sd = syn.synthetic_flow_state_dict(cfg, 1234)
    hsd = syn.synthetic_hifigan_state_dict(syn.HIFIGAN_COVOMIX, 1234)
    _, ids, cond, y0, _ = make_inputs(wl)
    with torch.inference_mode():
        t0 = time.perf_counter()
        mel = orc.flow_sample(sd, cfg, ids, cond, y0, cond_scale=0.7, method="euler", step_size=1.0 / wl["n_steps"])
        t1 = time.perf_counter()
        orc.hifigan_forward(hsd, syn.HIFIGAN_COVOMIX, mel[:, wl["prompt"]:, :].permute(0, 2, 1).contiguous())
        t2 = time.perf_counter()
    audio_s = (wl["N"] - wl["prompt"]) / FRAME_RATE
    return {"workload": wl["name"], "value": audio_s / (t2 - t0), "unit": "audio-s/s", "seconds": t2 - t0, "flow_seconds": t1 - t0,
            "vocoder_seconds": t2 - t1, "cores": cores, "kind": "port", "sample": "the whole job, measured once (no extrapolation)"}


def cpu_baseline(wl, cfg):
    from covomix_b200 import synthetic as syn
    from oracle import covomix_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = syn.synthetic_flow_state_dict(cfg, 1234)
    hsd = syn.synthetic_hifigan_state_dict(syn.HIFIGAN_COVOMIX, 1234)
    _, ids, cond, y0, _ = make_inputs(dict(wl, B=1))
    gen_frames = wl["N"] - wl["prompt"]
    mel = syn.synthetic_logmel(torch.Generator().manual_seed(31), 1, 80, gen_frames)
    n_eval = wl["n_steps"] * (2 if wl["method"] == "midpoint" else 1)
    with torch.inference_mode():
        orc.velocity_cfg(sd, cfg, y0, ids, cond, torch.tensor(0.5), 0.7)      # warm-up
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            orc.velocity_cfg(sd, cfg, y0, ids, cond, torch.tensor(0.5), 0.7)
        tv = (time.perf_counter() - t0) / reps
    orc.hifigan_forward(hsd, syn.HIFIGAN_COVOMIX, mel)
    t0 = time.perf_counter()
    orc.hifigan_forward(hsd, syn.HIFIGAN_COVOMIX, mel)
    th = time.perf_counter() - t0
    sec = wl["B"] * (n_eval * tv + th)
    extra = ""
    if "t2s" in wl:
        tt = cpu_t2s_step_seconds(wl)
        sec += wl["B"] * wl["t2s"]["steps"] * tt
        extra = f" + {tt * 1e3:.1f} ms per text-to-semantic decoding step (oracle, B = 1, 64 cached positions) x {wl['t2s']['steps']} steps"
    return {"value": wl["B"] * gen_frames / FRAME_RATE / sec, "unit": "audio-s/s", "cores": cores, "kind": "port",
            "c2_measured": cpu_c2_measured(),
            "sample": f"1 of {wl['B']} items, 1 of {n_eval} CFG velocity evaluations ({tv:.2f} s, fp32 torch CPU) + vocoder on 1 item "
                      f"({th:.2f} s){extra}; extrapolated linearly to the full step",
            "ode_step_ms_b1": tv * 1e3}


def cpu_t2s_step_seconds(wl):
    """Seconds per decoding step of the text-to-semantic oracle (B = 1, as the scripts run it) on the host cores:
    bounded sample of 48 steps after a 16-step warm-up (the reference's own step also re-projects the cross-attention
    context and re-embeds the prefix, so this is a lower bound for it)."""
    from covomix_b200 import synthetic as syn
    from oracle import t2s_oracle as t2o
    sd = syn.synthetic_t2s_state_dict(syn.COMIX, 1234)
    ids = syn.synthetic_text_ids(syn.COMIX, 1, wl["t2s"]["S"], seed=40, ragged=False)
    with torch.inference_mode():
        enc, mask = t2o.encode(sd, syn.COMIX, ids)
        st = t2o.DecoderState(sd, syn.COMIX, enc, mask)
        x = sd["start_token.speech"].expand(1, 1, -1)
        for _ in range(16):
            st.step(x)
        t0 = time.perf_counter()
        for _ in range(48):
            st.step(x)
        return (time.perf_counter() - t0) / 48


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=list(WORKLOADS) + ["c5"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="default workload only: skip the attached C2 / C4-pipelined / C5 measurements (`other_configs`)")
    args = ap.parse_args()
    if args.workload == "c5":
        run_vocoder_sweep(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
