#!/usr/bin/env python
"""bench.py -- the reference's headline workload on B200: one mean-teacher CRNN training step
(src/main.py:train_mt shapes: 12 synthetic + 12 real clips through the student forward + backward,
the 12 real clips through the teacher forward, BCE/MSE losses, Adam, EMA), plus the log-mel frontend.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).  `value` = student clips/s of the whole job with inputs resident in HBM;
`e2e` = the same step driven from pinned HOST buffers through the public trainer call (H2D of the
inputs and D2H of the losses inside the timed region); `roofline` = the dominant kernel class timed
live with CUDA events; `cpu_baseline` = the oracle port of the reference step on the host cores.
--impl reference times that CPU port alone (the reference is 100% Python and cannot travel to the GPU
box; oracle/crnn.py restates its modules with the same torch.nn layers -- see DESIGN.md).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "CRNN mean-teacher train clips/s"
N_SYN = N_REAL = 12
CLIP_BYTES_FRONTEND = 320000 * 4 + 1255 * 128 * 4
FLOP_PER_CLIP_FWD = 3.684e9
STEP_FLOP = (24 * 3 + 12) * FLOP_PER_CLIP_FWD             # 309.5 GFLOP
# CRNN_fpn adds per clip: the shared stage twice (conv 3x3 + GLU on 313 + 156 frames, 77 MMAC), rnn_2 / rnn_4
# ((156 + 78) / 313 of the GRU's 154 MMAC) and the two 512 -> 256 merges (61 MMAC): + 0.507 GFLOP forward
FLOP_PER_CLIP_FWD_FPN = FLOP_PER_CLIP_FWD + 0.507e9


class StdoutGuard:
    """Rank 0 must print exactly ONE line on stdout.  Libraries may write banners to the C-level stdout (NCCL prints its
    version there when NCCL_DEBUG is set on the box), so during the run file descriptor 1 points at stderr and the JSON
    line goes to the saved descriptor."""

    def __enter__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text):
        sys.stdout.flush()
        os.write(self.real, (text + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.real, 1)
        os.close(self.real)
        return False


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def cpu_mean_teacher(n_syn, n_real, steps, warmup, threads=None, fpn=False):
    """The oracle port of the reference step (torch.nn Conv2d / BatchNorm2d / GRU on the host cores)."""
    import torch
    from oracle import crnn as ocrnn
    from oracle import train as otrain
    from bsed_b200.utilities import synth
    # all host cores (torchrun exports OMP_NUM_THREADS=1; the baseline runs on rank 0 alone)
    torch.set_num_threads(threads or os.cpu_count() or 1)
    cls = ocrnn.OracleCRNNfpn if fpn else ocrnn.OracleCRNN
    oc = cls(**ocrnn.CRNN_KWARGS)
    op = ocrnn.OraclePredictor(**ocrnn.PREDICTOR_KWARGS)
    tc = cls(**ocrnn.CRNN_KWARGS)
    tp = ocrnn.OraclePredictor(**ocrnn.PREDICTOR_KWARGS)
    ocrnn.reference_style_init(oc, op, 1)
    ocrnn.reference_style_init(tc, tp, 2)
    for m in (oc, op, tc, tp):
        m.train()
    for prm in list(tc.parameters()) + list(tp.parameters()):
        prm.detach_()
    xs = torch.from_numpy(synth.make_logmel_like(n_syn, seed=3))
    xr = torch.from_numpy(synth.make_logmel_like(n_real, seed=4))
    ts = torch.from_numpy(synth.make_targets(n_syn, seed=5))
    opt = torch.optim.Adam(list(oc.parameters()) + list(op.parameters()), lr=5e-4, betas=(0.9, 0.999))
    for m in (oc, tc):
        m.set_dropout_keys(2023, 0, 0)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        otrain.mt_step(oc, op, tc, tp, opt, xr, xr, xs, ts, i, 500, ema_flavour="state_dict")
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return (n_syn + n_real) * steps / total, total / steps, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ns = nr = 2
    steps = max(1, min(args.steps, 20))
    warmup = max(1, min(args.warmup, 2))
    fpn = args.model == "crnn_fpn"
    cps, sec, threads = cpu_mean_teacher(ns, nr, steps, warmup, fpn=fpn)
    line = {
        "impl": "reference", "metric": METRIC, "value": cps, "unit": "clips/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "mean-teacher CRNN training step (main.py shapes), CPU sample of 2 synthetic + 2 real clips "
                               "per step instead of 12 + 12 (clips/s is per clip)", "model": args.model,
                   "parallelism": "cpu threads"},
        "cpu_baseline": {"value": cps, "unit": "clips/s", "cores": threads, "kind": "port",
                         "sample": f"{steps} steps x (2 syn + 2 real student clips, 2 teacher clips), oracle port of "
                                   "src/main.py:train_mt with the reference's torch.nn layers"},
        "e2e": {"value": cps, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_b200(args, out):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback of the product path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as ge
    if not os.path.exists(os.path.join(ROOT, "bird-sound-event-detecion_b200", "libbsed.so")):
        ge.build()
    from bsed_b200 import _lib, engine
    from bsed_b200.main import MeanTeacherTrainer
    from bsed_b200.models import CRNN, CRNN_fpn, Predictor
    from bsed_b200.utilities import synth
    from bsed_b200.utilities.utils import weights_init
    lib = _lib.load()
    peaks = load_peaks()
    fpn = args.model == "crnn_fpn"
    model_cls = CRNN_fpn if fpn else CRNN
    step_flop = (24 * 3 + 12) * (FLOP_PER_CLIP_FWD_FPN if fpn else FLOP_PER_CLIP_FWD)
    dev = torch.device("cuda", local)

    torch.manual_seed(2023 + rank)

    def make():
        m, p = model_cls(**engine.REFERENCE_CRNN_KWARGS), Predictor(**engine.REFERENCE_PREDICTOR_KWARGS)
        weights_init(m)
        weights_init(p)
        return m.to(dev).train(), p.to(dev).train()

    model, predictor = make()
    ema_model, ema_predictor = make()
    for prm in list(ema_model.parameters()) + list(ema_predictor.parameters()):
        prm.detach_()
    trainer = MeanTeacherTrainer(model, predictor, ema_model, ema_predictor, lr=5e-4, n_syn=N_SYN, n_real=N_REAL,
                                 dropout_seed=2023 + rank)

    # synthetic clips -> log-mel through our own frontend (also measured below)
    clips = torch.from_numpy(synth.make_clips(N_SYN + N_REAL, seed=2023 + rank)).to(dev)
    noise = torch.randn(N_REAL, 1255, 128, device=dev)
    mel = engine.melspec(clips)
    xs = engine.amp_to_db(mel[:N_SYN], 1255)[:, None].contiguous()
    x = engine.amp_to_db(mel[N_SYN:], 1255)[:, None].contiguous()
    x_ema = engine.amp_to_db(mel[N_SYN:], 1255, noise, 30.0)[:, None].contiguous()
    ts = torch.from_numpy(synth.make_targets(N_SYN, seed=7 + rank)).to(dev)
    rampup_len = 50 * 100

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # ---- device-resident run (value) with the dominant kernel class timed by CUDA events
    # host time to enqueue one step (no synchronisation inside): must stay below the device time or the GPU starves
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(3):
        trainer.step(x, x_ema, xs, ts, i, rampup_len)
    host_enqueue_ms = (time.perf_counter() - t0) / 3 * 1e3
    torch.cuda.synchronize()

    launches0 = lib.bsed_launch_count()
    lib.bsed_profile_begin(1)
    ms = timed(lambda i: trainer.step(x, x_ema, xs, ts, i, rampup_len), args.steps, args.warmup)
    import ctypes as C
    pm, pf, pb, pn = C.c_double(), C.c_double(), C.c_double(), C.c_int()
    _lib.check(lib.bsed_profile_end(C.byref(pm), C.byref(pf), C.byref(pb), C.byref(pn)), "profile_end")
    launches = (lib.bsed_launch_count() - launches0)
    launches_timed = launches * args.steps // (args.steps + args.warmup)
    value = (N_SYN + N_REAL) * world * args.steps / (ms * 1e-3)

    # ---- end to end: pinned host inputs -> H2D -> step -> D2H of the losses, every step.  The copies of step i+1 are
    # issued on a copy stream while step i computes (double-buffered device inputs), as a training loop's prefetcher
    # does; every step still moves its own inputs host -> device and its losses device -> host inside the timed region.
    hx, hxe, hxs, hts = [t.cpu().pin_memory() for t in (x, x_ema, xs, ts)]
    dbuf = [[torch.empty_like(t) for t in (x, x_ema, xs, ts)] for _ in range(2)]
    hloss = torch.empty(4).pin_memory()
    copy_stream = torch.cuda.Stream()
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    freed = [torch.cuda.Event(), torch.cuda.Event()]

    def h2d(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[slot])                     # the step that last read this slot is done
            for d, h in zip(dbuf[slot], (hx, hxe, hxs, hts)):
                d.copy_(h, non_blocking=True)
            ready[slot].record(copy_stream)

    for sl in (0, 1):
        freed[sl].record(torch.cuda.current_stream())
    e2e_state = {"primed": False}

    def e2e_step(i):
        slot = i & 1
        if not e2e_state["primed"]:
            h2d(slot)
            e2e_state["primed"] = True
        h2d(slot ^ 1)                                               # next step's inputs, overlapping this step
        torch.cuda.current_stream().wait_event(ready[slot])
        dx, dxe, dxs, dts = dbuf[slot]
        losses = trainer.step(dx, dxe, dxs, dts, 1000 + i, rampup_len)
        freed[slot].record(torch.cuda.current_stream())
        hloss.copy_(losses, non_blocking=True)
        torch.cuda.current_stream().synchronize()      # the caller reads the loss (reference: loss.item())

    ms_e2e = timed(e2e_step, args.steps, 1)
    e2e = (N_SYN + N_REAL) * world * args.steps / (ms_e2e * 1e-3)
    h2d_bytes = sum(t.numel() * 4 for t in (hx, hxe, hxs, hts))

    # ---- frontend: audio resident in HBM -> log-mel (second half of the metric)
    fe_clips = clips.repeat(16, 1)[:256].contiguous()                 # 256 clips = 328 MB > L2
    lib.bsed_profile_begin(5)
    ms_fe = timed(lambda i: engine.logmel(fe_clips, 1255), max(3, args.steps // 2), 3)       # STFT + mel + dB, one call
    fm = C.c_double()
    fn_ = C.c_int()
    _lib.check(lib.bsed_profile_end(C.byref(fm), None, None, C.byref(fn_)), "profile_end")
    fe_steps = max(3, args.steps // 2)
    fe_cps = 256 * world * fe_steps / (ms_fe * 1e-3)
    # the dB transform alone (the HBM-bound elementwise half of the frontend): algorithmic bytes = mel in + log-mel out
    fe_mel = engine.melspec(fe_clips)
    fe_out = torch.empty(256, 1255, 128, device=dev)
    lib.bsed_profile_begin(7)
    timed(lambda i: engine.amp_to_db(fe_mel, 1255, out=fe_out), fe_steps, 3)
    dm, db_, dn = C.c_double(), C.c_double(), C.c_int()
    _lib.check(lib.bsed_profile_end(C.byref(dm), None, C.byref(db_), C.byref(dn)), "profile_end")
    db_gbps = db_.value / (dm.value * 1e-3) / 1e9 if dm.value > 0 else None

    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)
        clocks = sampler.summary()
        conv_tflops = pf.value / (pm.value * 1e-3) / 1e12 if pm.value > 0 else None
        # the CPU port on this box's host cores, bounded sample
        if world == 1:
            cps_cpu, sec_cpu, threads = cpu_mean_teacher(2, 2, 12, 1, fpn=fpn)
            cpu_baseline = {"value": cps_cpu, "unit": "clips/s", "cores": threads, "kind": "port",
                            "sample": "12 steps of 2 synthetic + 2 real clips (oracle port of src/main.py:train_mt, "
                                      "torch.nn on host cores, %.1f s of CPU work)" % (sec_cpu * 12)}
        else:
            cpu_baseline = None     # reported at N = 1 only
        line = {
            "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32" if trainer.plan.precision == "tf32" else "f32",
            "data": "synthetic",
            "config": {"workload": "mean-teacher CRNN training step (src/main.py:train_mt, pretrain -mt): per GPU 12 synthetic + "
                                   "12 real clips student fwd+bwd, 12 clips teacher fwd (train mode), BCE+MSE, Adam lr 5e-4, "
                                   "state-dict EMA; log-mel features 1255x128 resident in HBM",
                       "precision": trainer.plan.precision + (" (tcgen05 kind::tf32 contractions, fp32 accumulate; everything "
                                                              "else fp32)" if trainer.plan.precision == "tf32" else ""),
                       "model": args.model + (" (src/models/CRNN.py:243-337)" if fpn else " (src/models/CRNN.py:178-240)"),
                       "clips_per_step_per_gpu": 24,
                       "parallelism": f"dp{world} (%s of the %.2f MB flat gradient)" % (
                           "one kernel per rank: reduce-scatter over NVLink peer memory + Adam + EMA + all-gather" if trainer.dp is not None
                           else "NCCL sum all-reduce", trainer.grads.numel() * 4 / 1e6),
                       "host_enqueue_ms_per_step": host_enqueue_ms,
                       "l2": "working set 2.7 GB of activations per step >> 126 MB L2 (no flush needed)",
                       "step_gflop_algorithmic": step_flop / 1e9},
            "e2e": {"value": e2e, "unit": "clips/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 16,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches_timed),
            "roofline": {"kernel": ("tc_conv_col_kernel + tc_kmajor_kernel (tcgen05 tf32 implicit-GEMM 3x3 conv forward + data "
                                    "gradient, TMA-fed; all such launches of the step)"
                                    if trainer.plan.precision == "tf32" else
                                    "gemm_nn_kernel<ConvRows> (implicit-GEMM 3x3 conv forward + data gradient, fp32 SIMT)"),
                         "bound": "tensor", "achieved": conv_tflops, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                         "frac": conv_tflops / peaks["tf_sustained"] if conv_tflops else None,
                         # ncu dram__bytes_read.sum + dram__bytes_write.sum summed over the 18 conv launches of one step,
                         # per launch (profiles/r01_ncu_tc_kernels.md); algorithmic bytes per launch beside it
                         "traffic": 3.899e7 if trainer.plan.precision == "tf32" else None,
                         "traffic_unit": "bytes per launch (average over the conv forward + data-gradient launches)",
                         "algorithmic_bytes_per_launch": pb.value / pn.value if pn.value else None,
                         "peak_source": peaks["source"] + " bf16 sustained (tf32 nominal dense peak is half of bf16)",
                         "launches": pn.value,
                         "share_of_step": pm.value / ms if ms else None,
                         "step_tflops": step_flop * args.steps / (ms * 1e-3) / 1e12},
            "frontend": {"metric": "log-mel frontend", "clips_per_s": fe_cps, "algorithmic_GBps": fe_cps * CLIP_BYTES_FRONTEND / 1e9,
                         "hbm_frac": fe_cps * CLIP_BYTES_FRONTEND / 1e9 / peaks["hbm"] / world,
                         "stft_mel_kernel_ms_per_256_clips": fm.value / max(1, fn_.value),
                         "stft_mel_fp32_tflops": 256 * 1255 * 70000.0 / (fm.value / max(1, fn_.value) * 1e-3) / 1e12,
                         "db_transform": {"bound": "hbm", "achieved_GBps": db_gbps, "peak_GBps": peaks["hbm"],
                                          "frac": db_gbps / peaks["hbm"] if db_gbps else None,
                                          "ms_per_256_clips": dm.value / max(1, dn.value),
                                          "note": "algorithmic bytes (amplitude-mel read once + log-mel written once); the per-clip "
                                                  "80 dB clamp needs the clip maximum first, so the kernel pair reads the mel twice"}},
            "cpu_baseline": cpu_baseline,
            "clocks": clocks,
        }
        out.emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_pseudo_label(args, out):
    """Secondary workload (BASELINE.json configs[4]): 1 h of synthetic audio (360 clips) -> log-mel -> CRNN + Predictor
    eval -> weak labels + strong events, clips sharded over the ranks, audio in pinned host memory."""
    import numpy as np
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from bsed_b200 import engine
    from bsed_b200.models import CRNN, Predictor
    from bsed_b200.pseudo_labeling import pseudo_label_stream
    from bsed_b200.utilities import synth
    from bsed_b200.utilities.utils import weights_init
    torch.manual_seed(2023)
    dev = torch.device("cuda", local)
    m, p = CRNN(**engine.REFERENCE_CRNN_KWARGS), Predictor(**engine.REFERENCE_PREDICTOR_KWARGS)
    weights_init(m)
    weights_init(p)
    m, p = m.to(dev).eval(), p.to(dev).eval()
    base = synth.make_clips(24, seed=11).reshape(-1)
    audio = torch.from_numpy(np.tile(base, 15)).pin_memory()          # 360 clips = 1 h at 32 kHz
    n_clips = audio.shape[0] // 320000
    for _ in range(max(1, args.warmup)):
        pseudo_label_stream(audio, m, p, batch_clips=48, rank=rank, world=world, gather=False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = max(1, args.steps // 4)
    for _ in range(reps):
        res = pseudo_label_stream(audio, m, p, batch_clips=48, rank=rank, world=world, gather=False)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        cps = n_clips * reps / (float(ms) * 1e-3)
        out.emit(json.dumps({"metric": "log-mel + CRNN pseudo-label inference clips/s", "value": cps, "unit": "clips/s",
                          "n_gpus": world, "steps": reps, "warmup": args.warmup, "ms_per_step": float(ms) / reps,
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                          "dtype": m.precision or engine.default_precision(), "data": "synthetic",
                          "config": {"workload": "1 h synthetic audio (360 x 10 s clips) from pinned host memory -> framed STFT "
                                                 "-> mel -> dB -> CRNN + Predictor eval -> weak labels + median-filtered events "
                                                 "(pseudo_labeling.pseudo_label_stream), clips sharded over ranks",
                                     "events_rank0": len(res["events"]), "audio_GBps": cps * 1280000 / 1e9}}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="crnn", choices=["crnn", "crnn_fpn"],
                    help="crnn = src/models/CRNN.py:CRNN (default, the headline); crnn_fpn = CRNN_fpn (SURVEY 8f-1)")
    ap.add_argument("--workload", default="train", choices=["train", "pseudo_label"],
                    help="train = the headline mean-teacher step (default); pseudo_label = configs[4] inference pipeline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return
    with StdoutGuard() as out:
        if args.workload == "pseudo_label":
            run_pseudo_label(args, out)
        else:
            run_b200(args, out)


if __name__ == "__main__":
    main()
