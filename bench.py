#!/usr/bin/env python3
"""Benchmark of the hot path (BASELINE.json): 1 s clips/sec for augment + log-mel + CNN+LSTM score.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload score|logmel|stream]

One "step" = one pass of the fused scoring path over one batch of synthetic clips (config 3 of
BASELINE.json: full augment (noise mix + shift + gain + 0.8-1.2x speed) + log-mel + CNN+LSTM scoring,
65,536 clips PER GPU -- weak scaling, clips are batch-sharded with no collective on the scoring path).
N > 1 is launched by torchrun, one rank per GPU; timing is CUDA events on the launching stream between
barrier+synchronize pairs, max over ranks.  Rank 0 prints ONE JSON line.

`value`  : device-resident throughput (inputs already in HBM when the timed region starts)
`e2e`    : same metric through the host-buffer C-ABI entry (ww_score_host): pinned-host H2D of the
           clips and augmentation parameters and D2H of logits/prob/decision inside the timed region
`roofline`: the dominant kernel (conv3 + ReLU + global mean) timed live with CUDA events inside the
           library (ww_profile), algorithmic FLOPs (conv3 = 377.5 MFLOP/clip) / time
           vs the measured bf16 dense peak in MEASURED_PEAKS.json
`cpu_baseline`: the oracle port (numpy log-mel per clip, as the reference does, + torch-CPU model)
           timed on this box's host cores on a bounded sample (rank 0, N = 1 only)
`--impl reference`: the reference's CPU path (oracle port; the reference is pure Python and cannot
           travel to the GPU box) with all host threads, same metric/config, bounded sample per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

CLIPS_PER_GPU = 65536
N_SAMPLES = 16000
FLOP_PER_CLIP = {"conv3": 2 * 2560 * 128 * 576, "conv2": 2 * 2560 * 64 * 288, "conv1": 2 * 2560 * 32 * 9,
                 "total": 475.5e6}
METRIC = "clips_per_sec_augment_logmel_cnn_lstm_score"
DTYPE_NAMES = {"fp32": "f32", "split2": "fp16 activations x (fp16 hi + e4m3 lo) weights on tcgen05, fp32 accumulate; fp32 front-end and head", "fp16": "fp16 (fp32 accumulate)"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons, pw = [], None, set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2]); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "power_w_max": max(pw) if pw else None, "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# synthetic inputs: the reference's create_sample_data recipe (wakeword_training_script.py:350-393)
def synth_clips_device(n, device, seed):
    g = torch.Generator(device=device).manual_seed(seed)
    t = torch.linspace(0, 1, N_SAMPLES, device=device)
    tone = 0.3 * torch.sin(2 * np.pi * 200 * t) + 0.2 * torch.sin(2 * np.pi * 400 * t)
    out = torch.empty((n, N_SAMPLES), device=device, dtype=torch.float32)
    step = 4096
    for i in range(0, n, step):
        m = min(step, n - i)
        z = torch.randn((m, N_SAMPLES), device=device, generator=g)
        pos = (torch.arange(i, i + m, device=device) % 3 == 0).float()[:, None]
        out[i:i + m] = pos * (0.1 * z + tone) + (1 - pos) * (0.2 * z)
    # The reference's recipe stores its clips as 16-bit WAV (sf.write, wakeword_training_script.py:371,380) and
    # reads them back with librosa.load (s / 32768): quantise the same way, so the fp32 device-resident arm and the
    # int16-PCM host arm score bit-identical samples.
    pcm = torch.clamp(torch.round(out * 32768.0), -32768, 32767).to(torch.int16)
    return pcm.to(torch.float32) / 32768.0, pcm


def seeded_state_dict(hidden=256, layers=2, n_classes=2, seed=0):
    """Random-init weights of the reference architecture, default-init shaped (U(-1/sqrt(fan_in), 1/sqrt(fan_in)), SURVEY.md
    appendix C) from a numpy PCG64 stream.  Same draws as the oracle's recipe, restated here so that the measured arm
    imports nothing from oracle/."""
    rng = np.random.default_rng(seed)
    sd = {}

    def u(shape, fan_in):
        b = 1.0 / np.sqrt(fan_in)
        return rng.uniform(-b, b, size=shape).astype(np.float32)

    for name, cout, cin in (("conv1", 32, 1), ("conv2", 64, 32), ("conv3", 128, 64)):
        sd[f"{name}.weight"] = u((cout, cin, 3, 3), cin * 9)
        sd[f"{name}.bias"] = u((cout,), cin * 9)
    for layer in range(layers):
        in_sz = 128 if layer == 0 else hidden
        sd[f"lstm.weight_ih_l{layer}"] = u((4 * hidden, in_sz), hidden)
        sd[f"lstm.weight_hh_l{layer}"] = u((4 * hidden, hidden), hidden)
        sd[f"lstm.bias_ih_l{layer}"] = u((4 * hidden,), hidden)
        sd[f"lstm.bias_hh_l{layer}"] = u((4 * hidden,), hidden)
    sd["fc.weight"] = u((n_classes, hidden), hidden)
    sd["fc.bias"] = u((n_classes,), hidden)
    return sd


def make_noise_bank(m=20, length=5 * 16000, seed=4321):
    """20 x 80,000 samples of 0.1 N(0,1) (the reference's background recipe, wakeword_training_script.py:382-388)."""
    return (0.1 * np.random.default_rng(seed).standard_normal((m, length))).astype(np.float32)


def draw_aug(n, seed):
    """Host draws, reference stage order (oracle.recipe.draw_aug_params restated with numpy for speed)."""
    from wakeword_jupyterlab_b200 import AugBatch, _lib
    r = np.random.default_rng(seed)
    flags = np.full(n, _lib.AUG_NORM_IN | _lib.AUG_NORM_OUT, np.uint32)
    do_shift, do_speed, do_noise = (r.random(n) < 0.8), (r.random(n) < 0.8), (r.random(n) < 0.8)
    shift = np.where(do_shift, (r.uniform(-0.3, 0.3, n) * 16000).astype(np.int32), 0).astype(np.int32)
    speed = r.integers(80, 121, n).astype(np.int32)
    do_speed &= speed != 100
    out_len = -(-100 * N_SAMPLES // speed)          # both already reduced by gcd-invariant ceil
    crop = np.where(do_speed & (out_len > N_SAMPLES), (r.random(n) * (np.maximum(out_len - N_SAMPLES, 0) + 1)).astype(np.int32), 0)
    flags |= np.where(do_shift, _lib.AUG_SHIFT, 0).astype(np.uint32)
    flags |= np.where(do_speed, _lib.AUG_SPEED, 0).astype(np.uint32)
    flags |= np.where(do_noise, _lib.AUG_NOISE, 0).astype(np.uint32)
    return AugBatch(flags, shift, np.where(do_speed, speed, 100).astype(np.int32), np.full(n, 100, np.int32),
                    crop.astype(np.int32), r.integers(0, 20, n).astype(np.int32),
                    r.integers(0, 5 * 16000 - N_SAMPLES + 1, n).astype(np.int32),
                    r.choice(np.array([0, 10, 20, 30, 40], np.float32), n).astype(np.float32),
                    np.ones(n, np.float32))


# ----------------------------------------------------------------------------------------------
# The reference's CPU path (oracle port: the reference is pure Python on librosa, which is neither installed nor able to
# travel to the GPU box, so the unmodified AudioProcessor.audio_to_mel cannot run here; oracle/ restates it and is pinned
# to the unmodified reference by tests/golden).  ONE protocol for `cpu_baseline` and `--impl reference`: one worker
# process per host core (1 thread each), every worker prepares its clips / noise bank / weights / augmentation draws ONCE
# (pool initializer) and a step times ONLY augment + log-mel + forward over the workers' clips.
CPU_CLIPS_PER_WORKER = 16
_W = {}


def _cpu_worker_init(seed0):
    import multiprocessing as mp
    from oracle import recipe as R
    torch.set_num_threads(1)
    ident = mp.current_process()._identity
    wid = ident[0] if ident else 0
    _W["clips"] = R.make_clips(CPU_CLIPS_PER_WORKER, seed=seed0 + wid)
    _W["bank"] = R.make_noise_bank()
    _W["aug"] = R.draw_aug_params(CPU_CLIPS_PER_WORKER, seed=seed0 + 7 * wid)
    _W["sd"] = {k: torch.from_numpy(v) for k, v in seeded_state_dict(256, seed=0).items()}


def _cpu_worker_step(_):
    """augment (per clip, like augment_audio) -> log-mel (per clip, like __getitem__) -> forward; returns busy seconds."""
    from oracle import augment as A, logmel as LM, model as M

    def run():
        aug = A.augment_batch(_W["clips"], _W["bank"], _W["aug"])
        feats = np.stack([LM.audio_to_mel(c) for c in aug]).astype(np.float32)[:, None]
        for i in range(0, len(feats), 32):                              # config 1: batches of up to 32
            M.forward_torch_cpu(torch.from_numpy(feats[i:i + 32]), _W["sd"])
    t0 = time.perf_counter()
    try:
        from threadpoolctl import threadpool_limits
        with threadpool_limits(limits=1):      # one BLAS/FFT thread per worker: no oversubscription
            run()
    except ImportError:
        run()
    return time.perf_counter() - t0


def cpu_pool_measure(steps, warmup):
    """-> (clips per step, mean seconds per step, worker processes)."""
    import multiprocessing as mp
    procs = max(1, min(os.cpu_count() or 1, 64))
    times = []
    with mp.get_context("fork").Pool(procs, initializer=_cpu_worker_init, initargs=(1000,)) as pool:
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            pool.map(_cpu_worker_step, range(procs), chunksize=1)
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
    return CPU_CLIPS_PER_WORKER * procs, float(np.mean(times)), procs


def cpu_sample_text(n, procs):
    return (f"{n} clips/step = {procs} worker processes x {CPU_CLIPS_PER_WORKER} clips, 1 thread each; inputs, noise bank, "
            "weights and augmentation draws prepared once per worker, the step times augment + log-mel + forward only; "
            "oracle port of the reference path (the unmodified librosa-based AudioProcessor cannot run on this box)")


def reference_arm(args, rank, world):
    if rank != 0:
        return
    n, sec, procs = cpu_pool_measure(args.steps, args.warmup)
    ms = 1e3 * sec
    v = n / sec
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "clips/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "config3: augment+logmel+cnn_lstm_score, CPU oracle port of the reference path",
                   "clips_per_step": n, "preset": "code (80x32, H=256)"},
        "cpu_baseline": {"value": v, "unit": "clips/s", "cores": procs, "kind": "port", "sample": cpu_sample_text(n, procs)},
        "e2e": {"value": v, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ----------------------------------------------------------------------------------------------
# BASELINE.json configs 2, 4 and 5.  Each returns a dict on every rank (rank 0 prints / embeds it); `--workload X` prints
# it as its own JSON line, the default run embeds short versions under "secondary" of the ONE contract line.
def measure_logmel(ww, dev, rank, world, steps, warmup, barrier, max_over_ranks, B=16384):
    """config 2: log-mel only, 16,384 clips per GPU (HBM roofline: 74,240 algorithmic bytes per clip)."""
    peaks = load_peaks()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng = ww.get_engine(device=dev.index)
    clips, _ = synth_clips_device(B, dev, seed=1234 + rank)
    out = torch.empty((B, 1, 80, eng.W), device=dev)
    for _ in range(warmup):
        eng.logmel(clips, normalize=False, out=out)
    barrier(); e0.record()
    for _ in range(steps):
        eng.logmel(clips, normalize=False, out=out)
    e1.record(); barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    gbs = 74240.0 * B / (ms * 1e-3) / 1e9
    return {"metric": "clips_per_sec_logmel_only", "value": B * world / (ms * 1e-3), "unit": "clips/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": "config2: log-mel only (80 mels, 80x32)",
                                            "clips_per_gpu": B, "l2": "input 1.05 GB > L2"},
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": gbs / peaks["hbm_gbs"], "traffic": None, "algorithmic_bytes_per_clip": 74240}}


def measure_train(ww, dev, rank, world, steps, warmup, barrier, max_over_ranks, B=4096):
    """config 5: CNN+LSTM training step (fwd + bwd + Adam) on on-GPU features, batch 4096 per GPU, gradient all-reduce
    over NCCL when world > 1 (dropout off: its masks are inputs of the kernels, not work)."""
    import torch.distributed as dist
    peaks = load_peaks()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    class MC(ww.ModelConfig):
        DROPOUT = 0.0
    net = ww.WakewordModel(MC).to(dev).train()
    net.load_state_dict({k: torch.from_numpy(v) for k, v in seeded_state_dict(256, seed=0).items()})
    tr = ww.WakewordTrainer(net, dev)
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    x = torch.randn((B, 1, 80, 32), device=dev, generator=g) * 15.0 - 40.0
    y = torch.randint(0, 2, (B,), device=dev, generator=g)
    for _ in range(warmup):
        tr.train_step(x, y)
    barrier(); e0.record()
    for _ in range(steps):
        loss, _ = tr.train_step(x, y)
    e1.record(); barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    # data-parallel replicas must hold identical weights after the all-reduced steps (they start identical and see
    # the same averaged gradient)
    in_sync = True
    if world > 1:
        chk = torch.stack([p.detach().double().sum() for p in net.parameters()])
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        in_sync = bool(torch.equal(lo, hi))
    flop = 1.419e9 * B
    # which backward kernels ran: the tensor-core ones (train_tc.cu) unless WW_TRAIN_KERNEL=fp32 / an fp32 context
    kern = "fp32" if os.environ.get("WW_TRAIN_KERNEL", "tc").startswith("f") else "tcgen05"
    train_dtype = "f32" if kern == "fp32" else "f16 hi+lo split products, f32 accumulate (weights, Adam, head: f32)"
    note = ("exact fp32 on CUDA cores (conv_fp32.cu)" if kern == "fp32" else
            "conv stack forward + weight / data gradients on tcgen05 (train_tc.cu); FLOP = 1.419 GFLOP per clip of useful work")
    return {"metric": "clips_per_sec_train_step", "value": B * world / (ms * 1e-3), "unit": "clips/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": train_dtype, "data": "synthetic",
            "config": {"workload": "config5: CNN+LSTM training step (fwd+bwd+Adam), on-GPU features",
                       "backward_kernels": kern,
                       "clips_per_gpu": B, "allreduce": "nccl sum of the flat fp32 gradient buffer" if world > 1 else "none",
                       "loss": float(loss.item()), "replicas_in_sync": in_sync},
            "roofline": {"bound": "tensor", "achieved": flop / (ms * 1e-3) / 1e12,
                         "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                         "frac": flop / (ms * 1e-3) / 1e12 / peaks["bf16_tflops_sustained"], "traffic": None,
                         "note": note}}


def measure_stream(ww, dev, rank, world, steps, warmup, barrier, max_over_ranks, conv_mode):
    """config 4: 1 h of 16 kHz audio, 1 s windows every 10 ms, each scored like predict_wakeword; the windows are split
    into `world` contiguous time ranges (strong scaling)."""
    from wakeword_jupyterlab_b200.sharding import window_shards
    peaks = load_peaks()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    T, N, hop = 57_600_000, N_SAMPLES, 160
    w0, n_win, s0, n_audio = window_shards(T, N, hop, world)[rank]
    net = ww.WakewordModel().to(dev).eval()
    net.load_state_dict({k: torch.from_numpy(v) for k, v in seeded_state_dict(256, seed=0).items()})
    net.conv_mode = conv_mode
    base = synth_clips_device(64, dev, seed=1234)[0].reshape(-1)                  # recipe audio, tiled over the hour
    audio = base.repeat((n_audio + base.numel() - 1) // base.numel())[:n_audio].contiguous()
    eng = net.engine()
    for _ in range(warmup):
        eng.score_stream(audio, hop)
    barrier(); e0.record()
    for _ in range(steps):
        prob1, dec = eng.score_stream(audio, hop)
    e1.record(); barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    total_win = 1 + (T - N) // hop
    tfs = total_win * FLOP_PER_CLIP["total"] / (ms * 1e-3) / 1e12 / world
    return {"metric": "windows_per_sec_streaming_1h_10ms_hop", "value": total_win / (ms * 1e-3),
            "unit": "windows/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": DTYPE_NAMES[conv_mode], "data": "synthetic",
            "config": {"workload": "config4: sliding-window detection over 1 h of 16 kHz audio",
                       "windows": total_win, "hop_samples": hop, "conv_mode": conv_mode,
                       "seconds_per_hour_of_audio": ms * 1e-3, "x_realtime": 3600.0 / (ms * 1e-3)},
            "roofline": {"bound": "tensor", "achieved": tfs, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                         "frac": tfs / peaks["bf16_tflops_sustained"], "traffic": None,
                         "note": "475.5 MFLOP per window (SURVEY 8d) per GPU"}}


def secondary_workload(args, rank, world, dev, conv_mode, barrier, max_over_ranks):
    """`--workload logmel|stream|train`: one JSON line for that configuration."""
    import wakeword_jupyterlab_b200 as ww
    if args.workload == "logmel":
        r = measure_logmel(ww, dev, rank, world, args.steps, args.warmup, barrier, max_over_ranks,
                           16384 if args.clips == CLIPS_PER_GPU else args.clips)
    elif args.workload == "train":
        r = measure_train(ww, dev, rank, world, args.steps, args.warmup, barrier, max_over_ranks,
                          4096 if args.clips == CLIPS_PER_GPU else args.clips)
    else:
        r = measure_stream(ww, dev, rank, world, args.steps, max(1, args.warmup - 2), barrier, max_over_ranks, conv_mode)
    if rank == 0:
        print(json.dumps(r))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--clips", type=int, default=CLIPS_PER_GPU, help="clips per GPU per step")
    ap.add_argument("--conv-mode", default=None)
    ap.add_argument("--workload", default="score", choices=["score", "logmel", "stream", "train"],
                    help="score = BASELINE config 3 (default, the contract line); logmel = config 2; stream = config 4; "
                         "train = config 5")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the short config 2 / 4 / 5 and strong-scaling legs")
    ap.add_argument("--e2e-fp32", action="store_true", help="e2e leg with fp32 host buffers instead of int16 PCM")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch.distributed as dist
    import wakeword_jupyterlab_b200 as ww
    from wakeword_jupyterlab_b200 import _lib, processor

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the first communicator is created: keep stdout for the ONE JSON
        # line by pointing fd 1 at stderr until that has happened
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    conv_mode = args.conv_mode or os.environ.get("WW_CONV_MODE", processor.DEFAULT_CONV_MODE)
    if args.workload != "score":
        secondary_workload(args, rank, world, dev, conv_mode, barrier, max_over_ranks)
        if world > 1:
            dist.destroy_process_group()
        return
    B = args.clips
    net = ww.WakewordModel().to(dev).eval()
    net.load_state_dict({k: torch.from_numpy(v) for k, v in seeded_state_dict(256, seed=0).items()})
    net.conv_mode = conv_mode
    eng = net.engine()
    bank = torch.from_numpy(make_noise_bank()).to(dev)

    def run_score(Bn, steps, warmup, want_e2e, sample_clocks):
        """One measurement of the scoring path on Bn clips per GPU: device-resident (CUDA events, max over ranks) and,
        optionally, end to end through ww_score_host_pcm16 from pinned host buffers."""
        clips, clips_pcm = synth_clips_device(Bn, dev, seed=1234 + rank)
        aug = draw_aug(Bn, seed=2024 + rank)
        aug_struct, keep = eng._aug_struct(aug, Bn)
        logits = torch.empty((Bn, 2), device=dev); prob1 = torch.empty((Bn,), device=dev)
        dec = torch.empty((Bn,), device=dev, dtype=torch.uint8)

        def step():
            eng.score_prepared(clips, aug_struct, bank, True, logits, prob1, dec)

        for _ in range(warmup):
            step()
        barrier()
        sampler = ClockSampler(local)
        if rank == 0 and sample_clocks:
            sampler.start()
        eng.profile(True)
        eng.profile_read(reset=True)
        l0 = eng.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        barrier()
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        r = {"B": Bn, "launches": eng.launches - l0, "stages": eng.profile_read(reset=True), "ms_step": ms_total / steps,
             "clocks": sampler.stop() if (rank == 0 and sample_clocks) else None, "e2e": None}
        eng.profile(False)
        if want_e2e:
            # host buffers hold the clips as they are on disk: int16 PCM (ww_score_host_pcm16; --e2e-fp32 for fp32 buffers),
            # in pinned memory placed on this GPU's NUMA node by the library (ww_host_alloc)
            h_clips = eng.host_buffer((Bn, N_SAMPLES), torch.float32 if args.e2e_fp32 else torch.int16)
            h_clips.copy_(clips if args.e2e_fp32 else clips_pcm)
            h_out = (eng.host_buffer((Bn, 2), torch.float32), eng.host_buffer((Bn,), torch.float32),
                     eng.host_buffer((Bn,), torch.uint8))
            for _ in range(2):
                eng.score_host(h_clips, aug=aug, noise_bank=bank, normalize=True, out=h_out)
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                eng.score_host(h_clips, aug=aug, noise_bank=bank, normalize=True, out=h_out)
            barrier()
            e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / steps
            assert torch.equal(h_out[2].to(dev), dec), "e2e decisions differ from the device-resident run"
            # what the box can deliver: every rank copies the same pinned buffer H2D at the same time, nothing else running
            d_tmp = torch.empty_like(h_clips, device=dev)
            d_tmp.copy_(h_clips, non_blocking=True)
            barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                d_tmp.copy_(h_clips, non_blocking=True)
            barrier()
            copy_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / 3
            nbytes = Bn * N_SAMPLES * h_clips.element_size()
            r["e2e"] = {"value": Bn * world / (e2e_ms / 1e3), "unit": "clips/s", "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": int(nbytes + Bn * 9 * 4), "d2h_bytes_per_step": int(Bn * (2 * 4 + 4 + 1)),
                        "host_input": "fp32 clips" if args.e2e_fp32 else "int16 PCM clips (as stored in the reference's WAV files)",
                        "h2d_ceiling_gbs": nbytes * world / (copy_ms * 1e-3) / 1e9,
                        "h2d_ceiling_ms_per_step": copy_ms,
                        "pinned": {"numa_node_of_gpu": getattr(h_clips, "ww_node", -1),
                                   "placement": {0: "plain pinned", 1: "mbind", 2: "first touch on a local CPU"}[getattr(h_clips, "ww_numa", 0)]}}
            del h_clips, h_out, d_tmp
        return r

    main_r = run_score(B, args.steps, args.warmup, not args.no_e2e, True)
    ms_step, launches, stages, clocks, e2e = main_r["ms_step"], main_r["launches"], main_r["stages"], main_r["clocks"], main_r["e2e"]
    value = B * world / (ms_step / 1e3)

    # ---- SURVEY 8e strong scaling: the SAME 65,536 clips split 65,536 / N per GPU (at N = 1 it is the line above)
    strong = None
    if not args.no_secondary:
        if world == 1:
            strong = {"global_clips": B, "clips_per_gpu": B, "value": value, "ms_per_step": ms_step,
                      "e2e_value": e2e["value"] if e2e else None, "note": "N = 1: identical to the weak-scaling line"}
        else:
            Bs = max(1, B // world)
            sr = run_score(Bs, args.steps, args.warmup, not args.no_e2e, False)
            strong = {"global_clips": Bs * world, "clips_per_gpu": Bs, "value": Bs * world / (sr["ms_step"] / 1e3),
                      "ms_per_step": sr["ms_step"], "e2e_value": sr["e2e"]["value"] if sr["e2e"] else None,
                      "e2e_ms_per_step": sr["e2e"]["ms_per_step"] if sr["e2e"] else None,
                      "h2d_ceiling_gbs": sr["e2e"]["h2d_ceiling_gbs"] if sr["e2e"] else None}

    # ---- BASELINE configs 2, 4, 5 in short form (every rank takes part; rank 0 embeds them)
    secondary = None
    if not args.no_secondary:
        secondary = {}
        for name, fn in (("logmel", lambda: measure_logmel(ww, dev, rank, world, 10, 3, barrier, max_over_ranks)),
                         ("stream", lambda: measure_stream(ww, dev, rank, world, 2, 1, barrier, max_over_ranks, conv_mode)),
                         ("train", lambda: measure_train(ww, dev, rank, world, 3, 1, barrier, max_over_ranks))):
            r = fn()
            secondary[name] = {k: r[k] for k in ("metric", "value", "unit", "ms_per_step", "steps", "warmup", "scaling", "config", "roofline") if k in r}
            torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    c3_ms, c3_n = stages["conv3"]
    per_launch_ms = c3_ms / max(c3_n, 1)
    clips_per_launch = B * args.steps / max(c3_n, 1)
    achieved = FLOP_PER_CLIP["conv3"] * clips_per_launch / (per_launch_ms * 1e-3) / 1e12 if c3_n else None
    peak = peaks["bf16_tflops_sustained"]
    traffic = None
    tp = os.path.join(ROOT, "profiles", "conv3_traffic_bytes_per_launch.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic = tj.get(conv_mode)
            if traffic is not None:     # measured on a launch of `clips_per_profiled_launch` clips; DRAM traffic is per clip
                traffic = traffic / tj.get("clips_per_profiled_launch", 4096) * clips_per_launch
        except Exception:
            traffic = None
    roof = {"bound": "tensor", "kernel": "conv3+relu+mean (" + conv_mode + ")", "achieved": achieved, "peak": peak,
            "unit": "TFLOP/s", "frac": (achieved / peak) if achieved else None, "traffic": traffic,
            "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']})",
            "avg_launch_ms": per_launch_ms, "launches_timed": c3_n, "clips_per_launch": clips_per_launch,
            "algorithmic_flop_per_clip": FLOP_PER_CLIP["conv3"],
            "whole_step_frac_of_peak": FLOP_PER_CLIP["total"] * B / (ms_step * 1e-3) / 1e12 / peak}
    stage_ms = {k: round(v[0] / args.steps, 4) for k, v in stages.items()}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        n_s, sec, procs = cpu_pool_measure(steps=3, warmup=1)           # same protocol as --impl reference
        cpu = {"value": n_s / sec, "unit": "clips/s", "cores": procs, "kind": "port", "sample": cpu_sample_text(n_s, procs)}

    out = {"metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": DTYPE_NAMES[conv_mode],
           "data": "synthetic",
           "config": {"workload": "config3: augment(noise mix+shift+gain+0.8-1.2x speed)+logmel+cnn_lstm_score",
                      "clips_per_gpu": B, "global_batch": B * world, "preset": "code (80x32, H=256)",
                      "conv_mode": conv_mode, "sharding": f"batch-sharded x{world}, no collective",
                      "l2": "inputs (4.2 GB/GPU) larger than L2; no explicit flush"},
           "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
           "stage_ms_per_step": stage_ms, "cpu_baseline": cpu, "strong": strong, "secondary": secondary}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
