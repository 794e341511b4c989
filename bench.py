#!/usr/bin/env python
"""Benchmark of the acoustic-frontend hot path (BASELINE.json metric: log-Mel audio-seconds per
second; HBM GB/s vs the measured peak; reference CPU path timed beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4|c5] [--impl reference]

One "step" = one pass of the hot path over one synthetic batch of the named workload.
  value  : audio-s/s with the waveforms already resident in HBM (CUDA events, max over ranks)
  e2e    : same metric through the public module API with PINNED HOST waveforms — the H2D copy
           of every step's batch and the D2H read of its result lengths are inside the timed region
  roofline / cpu_baseline : see DESIGN.md §5
  also   : (default workload only) the C1 kernel, the recipes' own blocks on the C2 batch — global statistics frozen
           (c2_recipe) and still averaged (c2_recipe_averaging) with SpecAugment's time warp — and C4 with its
           statistics all-reduce, each timed like `value`
Inputs cycle through a ring of distinct batches whose total footprint exceeds 3x the 126 MB L2,
so no step finds its input or output lines in cache.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

L2_BYTES = 126e6

WORKLOADS = {
    # BASELINE.json configs[0]
    "c1": dict(desc="C1: 16 kHz 80-Mel log10 Speech2MelSpec, 32 x 15 s (n_fft 512, win 400, hop 160)",
               sr=16000, B=32, fixed_len=240000, seed=1235,
               frontend=dict(n_mels=80, n_fft=512, win_length=0.025, hop_length=0.010, sr=16000),
               normalize=None, specaug=None),
    # BASELINE.json configs[1] — the configuration the metric is quoted on
    "c2": dict(desc="C2: ASR training batch, 64 x 2-20 s @16 kHz, 80-Mel (n_fft 512) + pre-emphasis 0.97 + "
                    "utterance norm + SpecAugment (2 freq / 2 time masks)",
               sr=16000, B=64, len_range=(32000, 320001), seed=1236,
               frontend=dict(n_mels=80, n_fft=512, win_length=0.025, hop_length=0.010, sr=16000, preemphasis=0.97),
               normalize=dict(norm_type="utterance"),
               specaug=dict(time_warp=False, freq_mask_width=30, freq_mask_num=2, time_mask_width=0.05,
                            time_mask_num=2)),
    # BASELINE.json configs[2]
    "c3": dict(desc="C3: TTS 22.05 kHz 80-Mel (n_fft 1024, hop 256, magnitude, ln, clamp 1e-5, energy), 128 x 10 s",
               sr=22050, B=128, fixed_len=220500, seed=1237,
               frontend=dict(sr=22050, n_mels=80, win_length=1024, hop_length=256, mag_spec=True, fmin=0, fmax=8000,
                             log_base=None, clamp=1e-5, return_energy=True),
               normalize=None, specaug=None),
    # BASELINE.json configs[3]: 2048 x 30 s split by utterance over 8 GPUs = 256 utterances per GPU (weak scaling
    # at other N), global normalisation in training mode: one packed (2D+1)-float all-reduce per step
    "c4": dict(desc="C4: sharded extraction, 256 x 30 s @16 kHz per GPU (2048 over 8), 80-Mel + global norm "
                    "(training, epoch=None) with the statistics all-reduce over NCCL",
               sr=16000, B=256, fixed_len=480000, seed=1238,
               frontend=dict(n_mels=80, n_fft=512, win_length=0.025, hop_length=0.010, sr=16000),
               normalize=dict(norm_type="global"), specaug=None, epoch=None),
    # BASELINE.json configs[4]: one batch of the offline TTS->ASR re-extraction (128 utterances 2-12 s @22.05 kHz):
    # linear magnitude spectrogram AND mel from one launch, mel -> global norm (epoch=1) -> SpecAugment
    "c5": dict(desc="C5: TTS->ASR chain batch, 128 x 2-12 s @22.05 kHz, linear (513 bins) + 80-Mel, global norm "
                    "(running average, epoch=1) + SpecAugment",
               sr=22050, B=128, len_range=(44100, 264601), seed=1239,
               frontend=dict(sr=22050, n_mels=80, win_length=1024, hop_length=256, mag_spec=True, fmin=0, fmax=8000,
                             log_base=None, clamp=1e-5),
               normalize=dict(norm_type="global"),
               specaug=dict(time_warp=False, freq_mask_width=30, freq_mask_num=2, time_mask_width=0.05,
                            time_mask_num=2), epoch=1, want_lin=True),
}


def workload_lengths(w, world=1, rank=0):
    """Lengths of this rank's utterances.  The global batch has B*world utterances sorted
    descending; rank r takes [world-1-r :: world] (iterator/abs.py:226-236) -> B per rank (weak scaling)."""
    n = w["B"] * world
    if "fixed_len" in w:
        lens = [w["fixed_len"]] * n
    else:
        g = torch.Generator().manual_seed(w["seed"])
        lo, hi = w["len_range"]
        lens = sorted(torch.randint(lo, hi, (n,), generator=g).tolist(), reverse=True)
        lens[0] = hi - 1  # the padded length is the nominal maximum
    return lens[world - 1 - rank::world]


def make_batch(w, lens, seed, device):
    """(B, T_max, 1) fp32 = 0.1*randn, zero beyond each length."""
    g = torch.Generator(device=device).manual_seed(seed)
    T = max(lens)
    wav = 0.1 * torch.randn(len(lens), T, 1, generator=g, device=device)
    idx = torch.arange(T, device=device)[None, :, None]
    wav.masked_fill_(idx >= torch.tensor(lens, device=device)[:, None, None], 0.0)
    return wav


class ClockSampler:
    """SM clock / throttle reasons sampled through NVML while the timed regions run."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.nv = None
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
# CPU baseline: the reference's op sequence on torch CPU tensors (oracle/ref_port_torch.py)
# ----------------------------------------------------------------------------------------------
def cpu_pass(w, wav, wl, weight, port):
    fc = w["frontend"]
    sr = fc.get("sr", 16000)
    hop = int(fc["hop_length"] * sr) if isinstance(fc["hop_length"], float) else fc["hop_length"]
    win = int(fc["win_length"] * sr) if isinstance(fc["win_length"], float) else fc["win_length"]
    n_fft = fc.get("n_fft") or win
    spec, fl, en = port.speech2linear(wav, wl, n_fft=n_fft, hop=hop, win=win, preemphasis=fc.get("preemphasis"),
                                      mag_spec=fc.get("mag_spec", False), return_energy=fc.get("return_energy", False))
    feat = port.linear2mel(spec, weight, clamp=fc.get("clamp", 1e-10), logging=True, log_base=fc.get("log_base", 10.0))
    if w["normalize"] is not None:
        if w["normalize"].get("norm_type") == "global":
            feat = port.norm_global_train(feat, fl, w.setdefault("_cpu_state", {}), w.get("epoch"))
        else:
            feat = port.norm_utterance(feat, fl)
    if w["specaug"] is not None:
        s = w["specaug"]
        feat = port.specaug(feat, fl, freq_mask_width=(0, s["freq_mask_width"]), freq_mask_num=s["freq_mask_num"],
                            time_mask_width=(0, s["time_mask_width"]), time_mask_num=s["time_mask_num"])
    return feat


def cpu_impl():
    """The CPU arm: the reference's OWN modules when `oracle/build_ref.py` has staged them under oracle/_ref
    (kind "reference"), else the bit-identical port of their op sequence (kind "port")."""
    try:
        from oracle import ref_modules  # imports oracle/_ref/speechain/module/...
        if ref_modules.available():
            return ref_modules, "reference"
    except Exception:  # noqa: BLE001
        pass
    from oracle import ref_port_torch as port
    return port, "port"


def time_cpu(w, steps, warmup, n_utt=None):
    """Time the CPU arm on the first `n_utt` utterances of the workload's batch (all by default)."""
    port, _ = cpu_impl()
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    lens = workload_lengths(w)
    if n_utt is not None:
        # keep the longest utterance (the padded length) and an evenly spaced subset of the others
        idx = sorted(set(np.linspace(0, len(lens) - 1, max(1, n_utt)).round().astype(int).tolist()))
        lens = [lens[i] for i in idx]
    wav = make_batch(w, lens, w["seed"], "cpu")
    wl = torch.tensor(lens, dtype=torch.int64)
    fc = w["frontend"]
    sr = fc.get("sr", 16000)
    win = int(fc["win_length"] * sr) if isinstance(fc["win_length"], float) else fc["win_length"]
    weight = port.mel_weight(sr, fc.get("n_fft") or win, fc["n_mels"], fc.get("fmin", 0.0), fc.get("fmax"))
    torch.manual_seed(1234)
    for _ in range(warmup):
        cpu_pass(w, wav.clone(), wl.clone(), weight, port)
    ts = []
    for _ in range(steps):
        x, l = wav.clone(), wl.clone()
        t0 = time.perf_counter()
        cpu_pass(w, x, l, weight, port)
        ts.append(time.perf_counter() - t0)
    audio_s = sum(lens) / w["sr"]
    return audio_s, ts, threads


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference(args, w, rank, world):
    if rank != 0:
        return
    # every step is one pass over a bounded sample of the batch, sized so that steps+warmup end in ~2.5 minutes
    B = w["B"]
    _, probe, _ = time_cpu(w, 1, 1)
    budget = 150.0
    n_utt = B
    if probe[0] * (args.steps + max(args.warmup, 1)) > budget:
        n_utt = int(max(1, min(B, B * budget / (probe[0] * (args.steps + max(args.warmup, 1))))))
    audio_s, ts, threads = time_cpu(w, args.steps, max(args.warmup, 1), n_utt=None if n_utt == B else n_utt)
    total = sum(ts)
    val = audio_s * len(ts) / total
    line = {
        "impl": "reference", "metric": "log-Mel audio-seconds per second", "value": val, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 1),
        "ms_per_step": 1e3 * total / len(ts), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "note": ("the reference's own modules (oracle/_ref, staged by oracle/build_ref.py) on "
                   "torch CPU" if cpu_impl()[1] == "reference" else "reference op sequence on torch CPU "
                   "(oracle/ref_port_torch.py); oracle/_ref is not staged on this box")},
        "cpu_baseline": {"value": val, "unit": "audio-s/s", "cores": threads, "kind": cpu_impl()[1],
                         "sample": f"{len(ts)} passes over {n_utt} of the batch's {B} utterances ({audio_s:.0f} audio-s "
                                   f"each), {cpu_model()}"},
        "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def measure_also(dev, world, rank, dist, peak, steps):
    """Two more configurations inside the same line: C1 (the one the 60 % target is quoted on: the fused kernel
    alone) and C4 (global normalisation in training mode: the path's only collective, the packed statistics
    all-reduce) — timed like the headline: CUDA events, ring of batches > 3x L2, max over ranks."""
    from speechain_b200.host import ops
    from speechain_b200.module.acoustic_frontend import AcousticFrontend
    out = {}

    def build(w):
        lens = workload_lengths(w, world, rank)
        chain = AcousticFrontend(frontend=dict(conf=w["frontend"]), normalize=w["normalize"], specaug=w["specaug"],
                                 distributed=world > 1).to(dev)
        chain.train()
        fe = chain.frontend
        B, T = len(lens), max(lens)
        hop, D = fe.speech2linear.hop_length, fe.output_size
        F = 1 + T // hop
        ring = max(2, int(np.ceil(3 * L2_BYTES / (4 * B * T + 4 * B * F * D))) + 1)
        batches = [make_batch(w, lens, w["seed"] + 100 * i + rank, dev) for i in range(ring)]
        wl_dev = torch.tensor(lens, dtype=torch.int64, device=dev)
        ops.HostLengths.remember(wl_dev, np.asarray(lens, dtype=np.int64))
        return chain, fe, lens, batches, wl_dev, ring, F, D

    def timed(fn, n):
        for i in range(3):
            fn(i)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / n

    # ---- C1: fused kernel alone (single-pass configuration: kernel == step) ----
    w = WORKLOADS["c1"]
    chain, fe, lens, batches, wl_dev, ring, F, D = build(w)
    ms = timed(lambda i: fe.run(batches[i % ring], wl_dev), max(50, min(steps, 300)))
    kb = 4 * sum(lens) + 4 * len(lens) * F * D
    out["c1"] = {"workload": w["desc"], "us_per_launch": ms * 1e3, "audio_s_per_s": world * sum(lens) / w["sr"] / (ms * 1e-3),
                 "roofline": {"bound": "hbm", "achieved": kb / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                              "frac": kb / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": kb}}
    del chain, fe, batches

    # ---- the recipes' own training step on the C2 batch: `normalize: True` = global statistics, frozen after
    #      max_epoch_num epochs (feat_norm.py:44-51), and the recipes' SpecAugment block, which leaves time_warp on
    #      (recipes/asr/librispeech/train-clean-100/exp_cfg/*.yaml:188-194) ----
    w = dict(WORKLOADS["c2"], normalize=dict(norm_type="global", max_epoch_num=1),
             specaug=dict(freq_mask_width=15, freq_mask_num=4, time_mask_width=0.05, time_mask_num=5))
    chain, fe, lens, batches, wl_dev, ring, F, D = build(w)
    nr = max(50, min(steps, 300))
    for i in range(3):
        chain(batches[i % ring], wl_dev, epoch=1)      # averaging epoch: fixes the statistics
    # the same blocks while the statistics are still being averaged (epochs 1 .. max_epoch_num): frontend + packed batch
    # statistics, one all-reduce of 2 D + 1 floats across the ranks, running average + normalise + warp + masks
    ms_avg = timed(lambda i: chain(batches[i % ring], wl_dev, epoch=1), nr)
    out["c2_recipe_averaging"] = {
        "workload": "C2 batch through the recipes' own blocks during epochs 1 .. max_epoch_num: global statistics still "
                    "averaged (all-reduced across the ranks every step) + SpecAugment with time warp",
        "ms_per_step": ms_avg, "value": world * sum(lens) / w["sr"] / (ms_avg * 1e-3), "unit": "audio-s/s", "steps": nr,
        "gpu_launches_per_step": 5}
    ms = timed(lambda i: chain(batches[i % ring], wl_dev, epoch=5), nr)
    t0 = time.perf_counter()
    for i in range(nr):
        chain(batches[i % ring], wl_dev, epoch=5)
    host_ms = (time.perf_counter() - t0) * 1e3 / nr
    torch.cuda.synchronize()
    out["c2_recipe"] = {"workload": "C2 batch through the recipes' own blocks: global statistics frozen (epoch > max_epoch_num) + "
                                    "SpecAugment with time warp, 4 frequency and 5 time masks: frontend kernel (normalised, "
                                    "frequency-masked rows) + in-place warp of rows [0, minlen)",
                        "ms_per_step": ms, "value": world * sum(lens) / w["sr"] / (ms * 1e-3), "unit": "audio-s/s",
                        "steps": nr, "host_enqueue_ms_per_step": host_ms, "gpu_launches_per_step": 2}
    del chain, fe, batches

    # ---- C4: frontend + global normalisation (training, epoch=None): statistics all-reduce every step ----
    w = WORKLOADS["c4"]
    chain, fe, lens, batches, wl_dev, ring, F, D = build(w)
    n4 = max(10, min(steps, 40))
    ms = timed(lambda i: chain(batches[i % ring], wl_dev, epoch=None), n4)
    c4 = {"workload": w["desc"], "ms_per_step": ms, "value": world * sum(lens) / w["sr"] / (ms * 1e-3), "unit": "audio-s/s",
          "steps": n4, "step_algorithmic_bytes": 4 * sum(lens) + 3 * 4 * len(lens) * F * D}
    c4["step_gbs"] = c4["step_algorithmic_bytes"] / (ms * 1e-3) / 1e9
    if dist is not None:
        # the collective alone, on the buffer the step exchanges
        packed = torch.zeros(2 * D + 1, dtype=torch.float32, device=dev)
        ar_ms = timed(lambda i: dist.all_reduce(packed), 50)
        c4["allreduce_us"] = ar_ms * 1e3
        # every rank must hold the same global statistics, bit for bit, and they must be what a single process
        # computes from the ranks' packed [sum of means | sum of stds | N] buffers
        torch.manual_seed(7)
        out_fe = fe.run(batches[0], wl_dev, want_stats=True, skip_pad=True)
        _, _, mine = ops.stats_finalize(out_fe["stats"], out_fe["feat_len"], chain.normalize.clamp, want_packed=True)
        chain(batches[0], wl_dev, epoch=None)
        gm, gs = chain.normalize.global_mean.clone(), chain.normalize.global_std.clone()
        allp = [torch.empty_like(mine) for _ in range(world)]
        allm = [torch.empty_like(gm) for _ in range(world)]
        alls = [torch.empty_like(gs) for _ in range(world)]
        dist.all_gather(allp, mine)
        dist.all_gather(allm, gm)
        dist.all_gather(alls, gs)
        same = all(torch.equal(allm[0], m) for m in allm) and all(torch.equal(alls[0], x) for x in alls)
        tot = torch.stack(allp).double().sum(0)
        ref_m, ref_s = (tot[:D] / tot[2 * D]).float(), (tot[D:2 * D] / tot[2 * D]).float()
        c4["ranks_agree_bitwise"] = bool(same)
        c4["max_abs_diff_vs_single_process_recompute"] = float(max((gm - ref_m).abs().max().item(),
                                                                   (gs - ref_s).abs().max().item()))
        c4["ranks"] = world
    else:
        c4["allreduce_us"] = None
        c4["ranks_agree_bitwise"] = None
        c4["ranks"] = 1
    out["c4"] = c4
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-also", action="store_true", help="skip the extra C1 / C4 measurements")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, w, rank, world)
        return

    from speechain_b200.host import ops
    from speechain_b200.module.acoustic_frontend import AcousticFrontend

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    if world > 1 and not os.environ.get("SFE_BENCH_NOBIND"):
        # one block of host cores per rank: eight interpreters that wander over the same cores (and their
        # hyper-thread siblings) were the scaling limit of round 1, not the GPUs
        try:
            cpus = sorted(os.sched_getaffinity(0))
            per = max(1, len(cpus) // world)
            os.sched_setaffinity(0, set(cpus[local * per:(local + 1) * per]))
        except (AttributeError, OSError):
            pass
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    steps, warmup = args.steps, max(args.warmup, 3)
    lens = workload_lengths(w, world, rank)
    B, T = len(lens), max(lens)
    audio_s = sum(lens) / w["sr"]

    chain = AcousticFrontend(frontend=dict(conf=w["frontend"]), normalize=w["normalize"], specaug=w["specaug"],
                             distributed=world > 1).to(dev)
    chain.train()
    fe = chain.frontend
    D, hop = fe.output_size, fe.speech2linear.hop_length
    F = 1 + T // hop
    flen = [n // hop + 1 for n in lens]
    out_bytes = 4 * B * F * D + (4 * B * F if fe.return_energy else 0)
    in_bytes_padded = 4 * B * T
    ring = max(2, int(np.ceil(3 * L2_BYTES / (in_bytes_padded + out_bytes))) + 1)
    batches = [make_batch(w, lens, w["seed"] + 100 * i + rank, dev) for i in range(ring)]
    wl_dev = torch.tensor(lens, dtype=torch.int64, device=dev)
    wl_host = torch.tensor(lens, dtype=torch.int64)
    ops.HostLengths.remember(wl_dev, wl_host.numpy())

    epoch = w.get("epoch")
    want_lin = w.get("want_lin", False)

    def step(i):
        if not want_lin:
            return chain(batches[i % ring], wl_dev, epoch=epoch)
        # C5: linear + mel from ONE launch, then the mel goes through normalise + SpecAugment in one pass
        out = fe.run(batches[i % ring], wl_dev, want_lin=True, want_stats=True)
        y = out["mel"]
        y._sfe_stats = (y._version, out["stats"])
        _, masks = chain.specaug.draw_device(B, D, out["feat_len_host"], dev)
        y, ylen = chain.normalize.forward(y, out["feat_len"], epoch=epoch, masks=masks, inplace=True)
        return y, ylen, out["lin"]

    torch.manual_seed(1234 + rank)
    sampler = ClockSampler(local)
    # spin the clocks up before anything is timed
    t_end = time.time() + (0.0 if os.environ.get("SFE_BENCH_NOSPIN") else 0.5)
    i = 0
    while time.time() < t_end:
        step(i)
        i += 1
        if i % 64 == 0:
            torch.cuda.synchronize()
    for i in range(warmup):
        step(i)
    barrier()
    sampler.start()

    # ---- (1) device-resident timing: exactly `steps` steps between two events ----
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    h0 = time.perf_counter()
    for i in range(steps):
        step(i)
    host_ms = (time.perf_counter() - h0) * 1e3   # time the host needed to enqueue the steps (no sync inside)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms, host_ms, -ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, host_ms_max, ms_min = float(t[0].item()), float(t[1].item()), -float(t[2].item())

    # ---- (2) the dominant kernel alone (fused frontend), same arguments as inside the step ----
    two_pass = w["normalize"] is not None
    kw = dict(want_stats=True, skip_pad=True) if two_pass else {}
    if want_lin:
        kw = dict(want_lin=True, want_stats=True)
    for i in range(3):
        fe.run(batches[i % ring], wl_dev, **kw)
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for i in range(steps):
        fe.run(batches[i % ring], wl_dev, **kw)
    k1.record()
    torch.cuda.synchronize()
    kern_ms = k0.elapsed_time(k1) / steps

    # ---- (3) end to end: pinned host waveforms -> H2D -> chain -> D2H of the result lengths ----
    host = [b.cpu().pin_memory() for b in batches[:2]]
    dbuf = [torch.empty_like(batches[0]) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]
    res_host = torch.empty(B, dtype=torch.int64).pin_memory()
    main_stream = torch.cuda.current_stream()

    def e2e_loop(n, upload=None):
        # double-buffered: the copy of step i+1 overlaps the kernels of step i
        if upload is None:
            upload = lambda k, slot: ops.upload_waveforms(host[k], wl_host, out=dbuf[slot])  # noqa: E731
        for s in range(2):
            free[s].record(main_stream)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(free[0])
            upload(0, 0)
            ready[0].record(copy_stream)
        for i in range(n):
            s = i % 2
            if i + 1 < n:
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(free[1 - s])
                    upload((i + 1) % 2, 1 - s)
                    ready[1 - s].record(copy_stream)
            main_stream.wait_event(ready[s])
            ylen = chain(dbuf[s], wl_host, epoch=epoch)[1]       # lengths arrive on the host, like a collated batch
            free[s].record(main_stream)
            res_host.copy_(ylen, non_blocking=True)
        main_stream.synchronize()

    e2e_steps = max(8, min(steps, 64))
    e2e_loop(4)
    barrier()
    t0 = time.perf_counter()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    e2e_loop(e2e_steps)
    g1.record()
    barrier()
    e2e_ms = g0.elapsed_time(g1)
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())

    # ---- (3b) the same loop with 16-bit PCM on the host (the corpus' own sample type): half the bytes on PCIe,
    #      the int16 -> float32 conversion of the data loader (soundfile) moves to the GPU ----
    host16 = [(h * 32768.0).round().clamp_(-32768, 32767).to(torch.int16).pin_memory() for h in host]
    stage16 = [torch.empty((ops.upload_i16_stage_bytes(B, T),), dtype=torch.uint8, device=dev) for _ in range(2)]
    up16 = lambda k, slot: ops.upload_waveforms_i16(host16[k], wl_host, out=dbuf[slot], stage=stage16[slot])  # noqa: E731
    e2e_loop(4, up16)
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    e2e_loop(e2e_steps, up16)
    g1.record()
    barrier()
    t = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e16_ms = float(t.item())
    del host16, stage16
    # ---- (3c) what the host -> device link gives a plain contiguous copy with all ranks copying at once: the ceiling
    #      of `e2e` on this box (the ranks share the host's memory system and, in pairs, its PCIe uplinks) ----
    probe_h = torch.empty(32 << 20, dtype=torch.float32).pin_memory()  # 128 MB
    probe_d = torch.empty(32 << 20, dtype=torch.float32, device=dev)
    probe_d.copy_(probe_h, non_blocking=True)
    h2d_plain_gbs = 0.0
    for _ in range(3):  # the best of three rounds: a ceiling
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(4):
            probe_d.copy_(probe_h, non_blocking=True)
        g1.record()
        barrier()
        t = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        h2d_plain_gbs = max(h2d_plain_gbs, 4 * probe_h.numel() * 4 / (float(t.item()) * 1e-3) / 1e9)
    del probe_h, probe_d
    also = None
    if not args.no_also and args.workload == "c2":
        del batches, host, dbuf
        torch.cuda.empty_cache()
        also = measure_also(dev, world, rank, dist, load_peaks()[0], steps)
    clocks = sampler.stop()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = load_peaks()
    S = sum(lens)
    valid_frames = sum(min(f, F) for f in flen)
    K = fe.speech2linear.output_size
    if want_lin:
        kern_bytes = 4 * S + 4 * B * F * K + 4 * B * F * D          # linear + raw mel written once
        step_bytes = kern_bytes + 2 * 4 * B * F * D                 # + mel read and re-written normalised/masked
    elif two_pass:
        kern_bytes = 4 * S + 4 * valid_frames * D          # reads valid samples, writes valid raw rows
        step_bytes = 4 * S + 4 * B * F * D                 # input once + padded output once
    else:
        kern_bytes = 4 * S + out_bytes
        step_bytes = kern_bytes
    achieved = kern_bytes / (kern_ms * 1e-3) / 1e9
    total_audio = audio_s * world
    line = {
        "metric": "log-Mel audio-seconds per second", "value": total_audio * steps / (ms_total * 1e-3),
        "unit": "audio-s/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_total / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "utterances_per_gpu": B, "audio_s_per_step_per_gpu": audio_s,
                   "padded_len": T, "frames": F, "sharding": "utterance stride [W-1-r::W], no data-path collective",
                   "cache": f"ring of {ring} distinct batches ({ring * (in_bytes_padded + out_bytes) / 1e6:.0f} MB) > 3x L2",
                   "value_note": "`value` feeds device-resident lengths whose host copy is already known (one "
                                 "HostLengths entry); a CUDA length tensor fresh from batch_to_cuda costs one .cpu() sync per "
                                 "step on top — `e2e` passes host lengths like a collated batch and pays everything"},
        "clocks": clocks,
        "host": {"enqueue_ms_per_step_max_over_ranks": host_ms_max / steps, "device_ms_per_step_min_over_ranks": ms_min / steps,
                 "cpus": os.cpu_count(), "affinity": len(os.sched_getaffinity(0))},
        "e2e": {"value": total_audio * e2e_steps / (e2e_ms * 1e-3), "unit": "audio-s/s",
                "h2d_bytes_per_step": ops.upload_bytes(wl_host) + 8 * B, "d2h_bytes_per_step": 8 * B, "steps": e2e_steps,
                "h2d_gbs_per_gpu": (ops.upload_bytes(wl_host) + 8 * B) * e2e_steps / (e2e_ms * 1e-3) / 1e9,
                "h2d_plain_copy_gbs_per_gpu": h2d_plain_gbs,
                "note": "pinned host (B,T,1) waveforms -> ops.upload_waveforms (valid samples; rows of similar length "
                        "as one pitched copy) -> public API "
                        "AcousticFrontend.forward; copy of step i+1 overlapped with step i; D2H = feature lengths "
                        "(the features stay on the GPU for the model); h2d_plain_copy_gbs_per_gpu = a plain 128 MB "
                        "pinned copy with all ranks copying at once (slowest rank, best of three rounds): the link's ceiling on this box"},
        "gpu_launches": steps * ({"utterance": 2, "global": 4}.get((w["normalize"] or {}).get("norm_type"), 1)),
        "roofline": {"bound": "hbm", "kernel": "sfe::fused_frontend_kernel", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": kern_bytes, "us_per_launch": kern_ms * 1e3,
                     "step_algorithmic_bytes": step_bytes,
                     "step_gbs": step_bytes / (ms_total / steps * 1e-3) / 1e9},
    }
    line["e2e_pcm16"] = {"value": total_audio * e2e_steps / (e2e16_ms * 1e-3), "unit": "audio-s/s",
                         "h2d_bytes_per_step": ops.upload_bytes(wl_host) // 2 + 8 * B, "d2h_bytes_per_step": 8 * B,
                         "note": "as e2e, but the pinned host batch holds 16-bit PCM (what the corpus files hold): "
                                 "ops.upload_waveforms_i16 moves two bytes per sample and converts x/32768 on the GPU, "
                                 "bit-identical to the float path (tests/test_gpu_parity.py)"}
    if also is not None:
        line["also"] = also
    gather = os.environ.get("SFE_KERNEL", "").startswith("g") or args.workload in ("c3", "c5")
    line["roofline"]["kernel"] = "sfe::fused_frontend_kernel" if gather else "sfe::tile_frontend_kernel"
    traffic_file = os.path.join(ROOT, "profiles", f"traffic_{args.workload}.json")
    if os.path.exists(traffic_file):
        with open(traffic_file) as f:
            line["roofline"]["traffic"] = json.load(f).get("dram_bytes_per_launch")
    if world == 1 and not args.no_cpu:
        # bounded sample: probe one pass, then as many full passes as fit in ~10 s of CPU work (3..60)
        _, probe, _ = time_cpu(w, 1, 1)
        n_pass = int(min(60, max(3, round(10.0 / max(probe[0], 1e-3)))))
        a_s, ts, threads = time_cpu(w, n_pass, 0)
        line["cpu_baseline"] = {"value": a_s * len(ts) / sum(ts), "unit": "audio-s/s", "cores": threads, "kind": cpu_impl()[1],
                                "sample": f"{len(ts)} full passes over one batch ({a_s:.0f} audio-s each) after 2 "
                                          f"warm-ups, {sum(ts):.1f} s of CPU work, {cpu_model()}"}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
