#!/usr/bin/env python
"""bench.py -- the reference's headline workload on B200: one mean-teacher CRNN training step
(src/main.py:train_mt shapes: 12 synthetic + 12 real clips through the student forward + backward,
the 12 real clips through the teacher forward, BCE/MSE losses, Adam, EMA), plus the log-mel frontend.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload train|ada|pseudo_label]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).
  value            student clips/s of the whole job, inputs resident in HBM, library DEFAULT precision (error-compensated
                   3xTF32 on the tcgen05 tensor cores: the <= 1e-3 parity mode)
  e2e              the same step driven from pinned HOST buffers through the public trainer call (H2D of the inputs
                   and D2H of the losses inside the timed region)
  roofline         the conv forward + data-gradient kernel class, timed live with CUDA events over the timed steps
  cpu_baseline     the oracle port of the reference step on the host cores, the SAME 12 + 12 + 12 configuration
  parity           (N = 1) the first GPU step against that CPU oracle step on the same clips / weights / dropout masks
  single_pass_tf32 the same step in the single-pass tf32 mode (what cuDNN gives the reference's convolutions on a GPU)
  gpu_eager_baseline  the reference's torch.nn modules (oracle/crnn.py restates them layer for layer) in eager mode on
                   the same B200, stock cuDNN / cuBLAS settings, same step -- the like-for-like GPU comparison
  sustained        the device-resident step again over >= 300 steps (>= 2 s of GPU time)
--impl reference times the CPU port alone (the reference is 100% Python and cannot travel to the GPU box;
oracle/crnn.py restates its modules with the same torch.nn layers -- see DESIGN.md).
"""
import argparse
import json
import math
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
# CRNN_fpn adds per clip: the shared stage twice (conv 3x3 + GLU on 313 + 156 frames, 77 MMAC), rnn_2 / rnn_4
# ((156 + 78) / 313 of the GRU's 154 MMAC) and the two 512 -> 256 merges (61 MMAC): + 0.507 GFLOP forward
FLOP_PER_CLIP_FWD_FPN = FLOP_PER_CLIP_FWD + 0.507e9
# conv forward + data gradient, algorithmic: blocks 1-6 forward 1.385 GMAC per clip (block 0 runs on
# the CUDA cores and is not in the class), data gradient of blocks 1-6 the same MACs
CONV_CLASS_GMAC_PER_CLIP = 1.385


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


def _conv_traffic_from_profile(launches_per_step, precision):
    """`roofline.traffic`: DRAM bytes (read + write) per launch of the conv-class kernels, averaged over the launches of one
    step, from the committed ncu capture of this build (`profiles/r02z_tc_kernels_metrics.csv`, command in
    `profiles/r02z_step_ncu.md`).  null unless that capture is of the default precision, holds exactly the launches the live
    run counted per step and the kernel is the same template (the capture does not travel into a different build)."""
    import collections
    import csv
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r02z_tc_kernels_metrics.csv")
    note = {"traffic": None, "traffic_note": "ncu dram bytes of these launches: profiles/r02z_step_ncu.md"}
    if precision != "tf32x3" or not os.path.exists(path):
        return note
    try:
        with open(path) as f:
            lines = [l for l in f if not l.startswith("==")]
        rows = collections.OrderedDict()
        for r in csv.DictReader(lines):
            if "tc_conv_col_kernel" not in r["Kernel Name"] or "dram__bytes" not in r["Metric Name"]:
                continue
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1.0)
            rows[int(r["ID"])] = rows.get(int(r["ID"]), 0.0) + float(r["Metric Value"].replace(",", "")) * scale
        if not rows or abs(len(rows) - launches_per_step) > 1e-6:
            return note
        return {"traffic": sum(rows.values()) / len(rows),
                "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the %d conv launches of one step "
                                "in profiles/r02z_tc_kernels_metrics.csv (ncu, same build; below the algorithmic bytes because the "
                                "teacher group and the data gradient find activations of the preceding launches in L2)" % len(rows)}
    except Exception:
        return note


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


# ------------------------------------------------------------------------------------------------
# baselines (test infrastructure: the oracle's restatement of the reference modules)
# ------------------------------------------------------------------------------------------------
def _oracle_models(fpn, seed_s=1, seed_t=2, dropout=0.5):
    import torch
    from oracle import crnn as ocrnn
    cls = ocrnn.OracleCRNNfpn if fpn else ocrnn.OracleCRNN
    kw = {**ocrnn.CRNN_KWARGS, "dropout": dropout}
    oc, op = cls(**kw), ocrnn.OraclePredictor(**ocrnn.PREDICTOR_KWARGS)
    tc, tp = cls(**kw), ocrnn.OraclePredictor(**ocrnn.PREDICTOR_KWARGS)
    ocrnn.reference_style_init(oc, op, seed_s, 0.2)
    ocrnn.reference_style_init(tc, tp, seed_t, 0.2)
    for m in (oc, op, tc, tp):
        m.train()
    for prm in list(tc.parameters()) + list(tp.parameters()):
        prm.detach_()
    return oc, op, tc, tp


def _step_inputs(n_syn, n_real):
    import torch
    from bsed_b200.utilities import synth
    xs = torch.from_numpy(synth.make_logmel_like(n_syn, seed=3))
    xr = torch.from_numpy(synth.make_logmel_like(n_real, seed=4))
    xr_ema = xr + 0.05 * torch.from_numpy(synth.make_logmel_like(n_real, seed=6))
    ts = torch.from_numpy(synth.make_targets(n_syn, seed=5))
    return xs, xr, xr_ema, ts


def _stock_dropout(mod):
    """The reference uses nn.Dropout; the hash dropout of the oracle is a parity-test device (numpy masks on the host)."""
    from torch import nn
    from oracle import crnn as ocrnn
    for name, child in list(mod.named_children()):
        if isinstance(child, (ocrnn.HashDropout, ocrnn.CycleHashDropout)):
            setattr(mod, name, nn.Dropout(child.p))
        else:
            _stock_dropout(child)


def cpu_mean_teacher(n_syn, n_real, steps, warmup, threads=None, fpn=False, keep_first=False):
    """The oracle port of the reference step (torch.nn Conv2d / BatchNorm2d / GRU / Dropout on the host cores), the full
    n_syn + n_real student clips and n_real teacher clips per step.  keep_first: one extra untimed step 0 with the hash
    dropout masks the CUDA kernels use; its losses / probabilities and the initial weights are returned so the GPU step
    can be compared with it (bench `parity`)."""
    import torch
    from oracle import train as otrain
    # all host cores (torchrun exports OMP_NUM_THREADS=1; the baseline runs on rank 0 alone)
    torch.set_num_threads(threads or os.cpu_count() or 1)
    oc, op, tc, tp = _oracle_models(fpn)
    xs, xr, xr_ema, ts = _step_inputs(n_syn, n_real)
    opt = torch.optim.Adam(list(oc.parameters()) + list(op.parameters()), lr=5e-4, betas=(0.9, 0.999))
    init = first = None
    if keep_first:
        init = [{k: v.clone() for k, v in m.state_dict().items()} for m in (oc, op, tc, tp)]

        def hook(tag):
            # device batch order: synthetic [0, ns), real [ns, ns + nr), teacher [ns + nr, ...)
            (tc if tag == "teacher" else oc).set_dropout_keys(2023, 0, {"teacher": n_syn + n_real, "syn": 0, "real": n_syn}[tag])
        loss, parts, outs = otrain.mt_step(oc, op, tc, tp, opt, xr, xr_ema, xs, ts, 0, 5000, ema_flavour="state_dict",
                                           dropout_hook=hook)
        first = dict(parts={k: float(v) for k, v in parts.items()}, strong=outs["strong"].numpy(), weak=outs["weak"].numpy(),
                     syn_strong=outs["syn_strong"].numpy())
    for m in (oc, tc):
        _stock_dropout(m)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        loss, _, _ = otrain.mt_step(oc, op, tc, tp, opt, xr, xr_ema, xs, ts, 1 + i, 5000, ema_flavour="state_dict")
        loss.item()              # src/main.py:511 reads the loss every iteration
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    res = dict(clips_per_s=(n_syn + n_real) * steps / total, sec_per_step=total / steps, threads=torch.get_num_threads())
    return (res, init, first) if keep_first else res


def gpu_eager_mean_teacher(n_syn, n_real, steps, warmup, fpn, dev, allow_tf32=None):
    """The reference's own module structure (torch.nn.Conv2d / BatchNorm2d / Linear / GRU / Dropout, restated layer for
    layer in oracle/crnn.py) in PyTorch eager mode on the B200, stock cuDNN / cuBLAS settings, driven by the reference's
    statement order (oracle/train.py:mt_step = src/main.py:250-254,339-343,376-477,517-523 incl. the per-step
    loss read-back of :511): the like-for-like GPU comparison BASELINE.md 5.4 asks for."""
    import torch
    from oracle import train as otrain
    if allow_tf32 is not None:
        torch.backends.cudnn.allow_tf32 = allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = allow_tf32
    oc, op, tc, tp = _oracle_models(fpn)

    for m in (oc, tc):
        _stock_dropout(m)
    for m in (oc, op, tc, tp):
        m.to(dev).train()
    xs, xr, xr_ema, ts = [t.to(dev) for t in _step_inputs(n_syn, n_real)]
    opt = torch.optim.Adam(list(oc.parameters()) + list(op.parameters()), lr=5e-4, betas=(0.9, 0.999))

    def step(i):
        loss, _, _ = otrain.mt_step(oc, op, tc, tp, opt, xr, xr_ema, xs, ts, i, 5000, ema_flavour="state_dict")
        return loss.item()       # src/main.py:511 reads the loss every iteration

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return dict(value=(n_syn + n_real) / (ms * 1e-3), unit="clips/s", ms_per_step=ms, steps=steps, warmup=warmup,
                cudnn_allow_tf32=bool(torch.backends.cudnn.allow_tf32),
                matmul_allow_tf32=bool(torch.backends.cuda.matmul.allow_tf32),
                kind="port (oracle/crnn.py restates src/models/{CNN,RNN,CRNN}.py with the same torch.nn layers; "
                     "oracle/train.py:mt_step is src/main.py:train_mt's statement order)",
                torch=torch.__version__, cudnn=torch.backends.cudnn.version())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fpn = args.model == "crnn_fpn"
    # the SAME configuration as the GPU arm: 12 synthetic + 12 real student clips, 12 teacher clips per step.  One step
    # is ~5 s of CPU work on 16 cores, so more than 40 steps would not end "within a few minutes": capped there.
    steps = max(1, min(args.steps, 40))
    warmup = max(0, min(args.warmup, 10))
    r = cpu_mean_teacher(N_SYN, N_REAL, steps, warmup, fpn=fpn)
    cps, sec, threads = r["clips_per_s"], r["sec_per_step"], r["threads"]
    line = {
        "impl": "reference", "metric": METRIC, "value": cps, "unit": "clips/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "mean-teacher CRNN training step (src/main.py:train_mt, pretrain -mt): 12 synthetic + 12 real "
                               "clips student fwd+bwd, 12 clips teacher fwd (train mode), BCE+MSE, Adam lr 5e-4, state-dict "
                               "EMA; log-mel features 1255x128 resident in host memory",
                   "model": args.model, "clips_per_step_per_gpu": 24, "parallelism": "cpu threads"},
        "cpu_baseline": {"value": cps, "unit": "clips/s", "cores": threads, "kind": "port",
                         "sample": f"{steps} steps x (12 syn + 12 real student clips, 12 teacher clips) after {warmup} warm-up "
                                   "steps, oracle port of src/main.py:train_mt with the reference's torch.nn layers"},
        "e2e": {"value": cps, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args, out):
    import ctypes as C
    import numpy as np
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
    default_precision = engine.default_precision()

    torch.manual_seed(2023 + rank)

    def make(precision=None):
        m, p = model_cls(**engine.REFERENCE_CRNN_KWARGS, precision=precision), Predictor(**engine.REFERENCE_PREDICTOR_KWARGS)
        weights_init(m)
        weights_init(p)
        return m.to(dev).train(), p.to(dev).train()

    def make_trainer(precision=None):
        model, predictor = make(precision)
        ema_model, ema_predictor = make(precision)
        for prm in list(ema_model.parameters()) + list(ema_predictor.parameters()):
            prm.detach_()
        return MeanTeacherTrainer(model, predictor, ema_model, ema_predictor, lr=5e-4, n_syn=N_SYN, n_real=N_REAL,
                                  dropout_seed=2023 + rank, precision=precision)

    trainer = make_trainer()

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

    def timed(fn, steps, warmup, before_timed=None):
        """W untimed warm-up steps, then EXACTLY `steps` steps between two events, barrier + synchronize on both sides,
        max over ranks.  before_timed runs between the warm-up and the timed region (profile / launch-counter reset)."""
        for i in range(warmup):
            fn(i)
        sync_all()
        if before_timed is not None:
            before_timed()
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

    # ---- device-resident run (value): the public trainer call, which replays ONE CUDA graph per iteration from its
    # second call on (main.py: MeanTeacherTrainer.step).  Host time to enqueue one step (no synchronisation inside) is
    # measured on the replay path.
    for i in range(3):
        trainer.step(x, x_ema, xs, ts, i, rampup_len)           # eager, capture, first replay
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(3, 13):
        trainer.step(x, x_ema, xs, ts, i, rampup_len)
    host_enqueue_ms = (time.perf_counter() - t0) / 10 * 1e3
    torch.cuda.synchronize()

    ms = timed(lambda i: trainer.step(x, x_ema, xs, ts, 13 + i, rampup_len), args.steps, args.warmup)
    value = (N_SYN + N_REAL) * world * args.steps / (ms * 1e-3)
    graphed = bool(trainer._graphs)
    launches_per_step = next(iter(trainer.graph_launches.values())) if graphed else None

    # ---- the conv kernel class timed by CUDA events: the same steps enqueued kernel by kernel (events cannot be read
    # out of a replayed graph), its share taken against THIS run's own step time
    counters = {}

    def begin_profile():
        counters["launches0"] = lib.bsed_launch_count()
        lib.bsed_profile_begin(1)

    trainer.use_graph = False
    ms_eager = timed(lambda i: trainer.step(x, x_ema, xs, ts, 100 + i, rampup_len), args.steps, args.warmup, begin_profile)
    launches_eager = lib.bsed_launch_count() - counters["launches0"]
    pm, pf, pb, pn = C.c_double(), C.c_double(), C.c_double(), C.c_int()
    _lib.check(lib.bsed_profile_end(C.byref(pm), C.byref(pf), C.byref(pb), C.byref(pn)), "profile_end")
    trainer.use_graph = graphed
    launches = launches_per_step * args.steps if graphed else launches_eager

    # ---- the same device-resident step over >= 300 steps (>= 2 s of GPU time)
    sus_steps = max(480, args.steps)          # >= 2 s of GPU time at 5.3 ms per step (4.8 ms single-pass)
    ms_sus = timed(lambda i: trainer.step(x, x_ema, xs, ts, 200 + i, rampup_len), sus_steps, 0)
    sustained = {"steps": sus_steps, "ms_per_step": ms_sus / sus_steps, "seconds": ms_sus * 1e-3,
                 "value": (N_SYN + N_REAL) * world * sus_steps / (ms_sus * 1e-3), "unit": "clips/s"}

    # ---- end to end: pinned host inputs -> H2D -> step -> D2H of the losses, every step.  The copies of step i+1 are
    # issued on a copy stream while step i computes (double-buffered device inputs), as a training loop's prefetcher
    # does; every step still moves its own inputs host -> device and its losses device -> host inside the timed region.
    hx, hxe, hxs, hts = [t.cpu().pin_memory() for t in (x, x_ema, xs, ts)]
    dbuf = [[torch.empty_like(t) for t in (x, x_ema, xs, ts)] for _ in range(2)]
    hloss = [torch.empty(4).pin_memory() for _ in range(2)]
    loss_done = [torch.cuda.Event(), torch.cuda.Event()]
    loss_log = []
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
        hloss[slot].copy_(losses, non_blocking=True)
        loss_done[slot].record(torch.cuda.current_stream())
        # the caller reads every step's loss on the host (reference: loss.item()), one step late: step i is already queued
        # when the host blocks on step i - 1, so the graph launch never waits for the host; the last step's loss is read
        # by the closing synchronisation of the timed region
        if i > 0:
            loss_done[slot ^ 1].synchronize()
            loss_log.append(float(hloss[slot ^ 1].sum()))

    ms_e2e = timed(e2e_step, args.steps, 2)
    torch.cuda.synchronize()
    loss_log.append(float(hloss[(args.steps - 1) & 1].sum()))
    assert all(math.isfinite(v) for v in loss_log), "e2e: non-finite loss read back"
    e2e = (N_SYN + N_REAL) * world * args.steps / (ms_e2e * 1e-3)
    h2d_bytes = sum(t.numel() * 4 for t in (hx, hxe, hxs, hts))

    # ---- the single-pass tf32 mode on the same step (not the parity mode; what cuDNN TF32 convolutions correspond to)
    other = "tf32" if default_precision != "tf32" else "tf32x3"
    tr2 = make_trainer(other)
    ms2 = timed(lambda i: tr2.step(x, x_ema, xs, ts, i, rampup_len), args.steps, args.warmup)
    # its conv class, timed the same way (kernel by kernel, CUDA events)
    tr2.use_graph = False
    ms2_eager = timed(lambda i: tr2.step(x, x_ema, xs, ts, 100 + i, rampup_len), args.steps, 3, lambda: lib.bsed_profile_begin(1))
    qm, qf = C.c_double(), C.c_double()
    _lib.check(lib.bsed_profile_end(C.byref(qm), C.byref(qf), None, None), "profile_end")
    conv2 = qf.value / (qm.value * 1e-3) / 1e12 if qm.value > 0 else None
    other_mode = {"precision": other, "value": (N_SYN + N_REAL) * world * args.steps / (ms2 * 1e-3), "unit": "clips/s",
                  "ms_per_step": ms2 / args.steps,
                  "roofline": {"kernel": "the same conv forward + data-gradient class", "achieved": conv2, "unit": "TFLOP/s",
                               "peak": peaks["tf_sustained"], "frac": conv2 / peaks["tf_sustained"] if conv2 else None,
                               "ms_per_step": qm.value / args.steps, "share_of_step": qm.value / ms2_eager if ms2_eager else None}}
    del tr2
    torch.cuda.empty_cache()

    # ---- frontend: audio resident in HBM -> log-mel (second half of the metric)
    fe_clips = clips.repeat(16, 1)[:256].contiguous()                 # 256 clips = 328 MB > L2
    fe_steps = max(3, args.steps // 2)
    ms_fe = timed(lambda i: engine.logmel(fe_clips, 1255), fe_steps, 3, lambda: lib.bsed_profile_begin(5))   # STFT + mel + dB
    fm = C.c_double()
    fn_ = C.c_int()
    _lib.check(lib.bsed_profile_end(C.byref(fm), None, None, C.byref(fn_)), "profile_end")
    fe_cps = 256 * world * fe_steps / (ms_fe * 1e-3)
    # the dB transform alone (the HBM-bound elementwise half of the frontend): algorithmic bytes = mel in + log-mel out
    fe_mel = engine.melspec(fe_clips)
    fe_out = torch.empty(256, 1255, 128, device=dev)
    timed(lambda i: engine.amp_to_db(fe_mel, 1255, out=fe_out), fe_steps, 3, lambda: lib.bsed_profile_begin(7))
    dm, db_, dn = C.c_double(), C.c_double(), C.c_int()
    _lib.check(lib.bsed_profile_end(C.byref(dm), None, C.byref(db_), C.byref(dn)), "profile_end")
    db_gbps = db_.value / (dm.value * 1e-3) / 1e9 if dm.value > 0 else None
    del fe_clips, fe_mel, fe_out
    torch.cuda.empty_cache()

    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)
        clocks = sampler.summary()
        conv_tflops = pf.value / (pm.value * 1e-3) / 1e12 if pm.value > 0 else None
        cpu_baseline = parity = eager = None
        if world == 1:
            # ---- the reference modules in PyTorch eager mode on this B200 (stock cuDNN / cuBLAS)
            eager = gpu_eager_mean_teacher(N_SYN, N_REAL, 20, 5, fpn, dev)
            torch.cuda.empty_cache()
            # ---- the CPU port on this box's host cores: the SAME 12 + 12 + 12 step, 3 steps after 1 warm-up, and the
            # first GPU step (default precision and single-pass tf32) against its first step
            r, init, first = cpu_mean_teacher(N_SYN, N_REAL, 3, 1, fpn=fpn, keep_first=True)
            cpu_baseline = {"value": r["clips_per_s"], "unit": "clips/s", "cores": r["threads"], "kind": "port",
                            "ms_per_step": r["sec_per_step"] * 1e3,
                            "sample": "3 steps (after 1 warm-up) of 12 synthetic + 12 real student clips and 12 teacher clips "
                                      "(oracle port of src/main.py:train_mt, torch.nn on host cores, %.0f s of CPU work)"
                                      % (r["sec_per_step"] * 4)}
            xs_c, xr_c, xr_ema_c, ts_c = _step_inputs(N_SYN, N_REAL)
            parity = {"against": "first step of the CPU oracle port on the same clips, weights and dropout masks "
                                 "(12 + 12 + 12 clips)", "tolerance_north_star": 1e-3}
            want = np.array([first["parts"][k] for k in ("strong_class", "weak_class", "cons_strong", "cons_weak")])
            for prec in (default_precision, other):
                mods = []
                for sd_c, sd_p in ((init[0], init[1]), (init[2], init[3])):
                    m_, p_ = model_cls(**engine.REFERENCE_CRNN_KWARGS, precision=prec), Predictor(**engine.REFERENCE_PREDICTOR_KWARGS)
                    m_.load_state_dict(sd_c)
                    p_.load_state_dict(sd_p)
                    mods += [m_.to(dev).train(), p_.to(dev).train()]
                for prm in list(mods[2].parameters()) + list(mods[3].parameters()):
                    prm.detach_()
                tr = MeanTeacherTrainer(*mods, lr=5e-4, n_syn=N_SYN, n_real=N_REAL, dropout_seed=2023, precision=prec)
                got = tr.step(xr_c.to(dev), xr_ema_c.to(dev), xs_c.to(dev), ts_c.to(dev), 0, 5000).cpu().numpy()
                parity[prec] = {
                    "strong_max_abs_err": float(np.abs(tr.last["strong"][N_SYN:].cpu().numpy() - first["strong"]).max()),
                    "syn_strong_max_abs_err": float(np.abs(tr.last["strong"][:N_SYN].cpu().numpy() - first["syn_strong"]).max()),
                    "weak_max_abs_err": float(np.abs(tr.last["weak"][N_SYN:].cpu().numpy() - first["weak"]).max()),
                    "loss_terms_max_rel_err": float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-6)))}
                del tr, mods
                torch.cuda.empty_cache()
        conv_flop_algorithmic = (36 + 24) * CONV_CLASS_GMAC_PER_CLIP * 2e9 if not fpn else None
        line = {
            "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"tf32x3": "tf32x3", "tf32": "tf32", "fp32": "f32"}[trainer.plan.precision],
            "data": "synthetic",
            "config": {"workload": "mean-teacher CRNN training step (src/main.py:train_mt, pretrain -mt): per GPU 12 synthetic + "
                                   "12 real clips student fwd+bwd, 12 clips teacher fwd (train mode), BCE+MSE, Adam lr 5e-4, "
                                   "state-dict EMA; log-mel features 1255x128 resident in HBM",
                       "precision": {"tf32x3": "tf32x3: error-compensated 3xTF32 on tcgen05 (a*w_hi + a*w_lo + a_lo*w_hi, fp32 "
                                               "accumulation in TMEM) for every forward / data-gradient contraction = fp32-grade, "
                                               "the <= 1e-3 parity mode; weight-gradient reductions single-pass tf32; everything "
                                               "else fp32",
                                     "tf32": "tf32 (single-pass tcgen05 kind::tf32 contractions, fp32 accumulate; everything else fp32)",
                                     "fp32": "fp32 CUDA cores"}[trainer.plan.precision],
                       "model": args.model + (" (src/models/CRNN.py:243-337)" if fpn else " (src/models/CRNN.py:178-240)"),
                       "clips_per_step_per_gpu": 24,
                       "parallelism": f"dp{world} (%s of the %.2f MB flat gradient)" % (
                           "one kernel per rank: reduce-scatter over NVLink peer memory + Adam + EMA + all-gather" if trainer.dp is not None
                           else "NCCL sum all-reduce", trainer.grads.numel() * 4 / 1e6),
                       "host_enqueue_ms_per_step": host_enqueue_ms,
                       "cuda_graph": ("one graph launch per iteration (device-resident step state: dropout keys, Adam bias "
                                      "corrections, EMA coefficient, consistency weight)" if graphed else "off"),
                       "ms_per_step_without_graph": ms_eager / args.steps,
                       "l2": "working set 2.7 GB of activations per step >> 126 MB L2 (no flush needed)",
                       "step_gflop_algorithmic": step_flop / 1e9},
            "e2e": {"value": e2e, "unit": "clips/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 16,
                    "ms_per_step": ms_e2e / args.steps,
                    "note": "every step: its inputs host -> device (copy stream, double-buffered, overlapping the previous step), "
                            "trainer.step, its four loss terms device -> host; the host reads each step's loss one step late"},
            "gpu_launches": int(launches),
            "sustained": sustained,
            "roofline": {"kernel": ("tc_conv_col_kernel + tc_kmajor_kernel conv launches (tcgen05 implicit-GEMM 3x3 conv forward "
                                    "+ data gradient of blocks 1-6, TMA-fed; all such launches of the timed steps)"
                                    if trainer.plan.precision != "fp32" else
                                    "gemm_nn_kernel<ConvRows> (implicit-GEMM 3x3 conv forward + data gradient, fp32 SIMT)"),
                         "bound": "tensor",
                         # ALGORITHMIC flops of the launches (2*9*Cin*Cout per output pixel; the 3xTF32 mode executes three
                         # MMAs per algorithmic one, the pixel-pair view of block 1 counts its true 16 input channels)
                         "achieved": conv_tflops, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                         "frac": conv_tflops / peaks["tf_sustained"] if conv_tflops else None,
                         "frac_of_executed_mmas": (conv_tflops / peaks["tf_sustained"] * (3 * 36 + 24) / 60.0
                                                   if conv_tflops and trainer.plan.precision == "tf32x3" else None),
                         **_conv_traffic_from_profile(pn.value / args.steps if args.steps else 0, trainer.plan.precision),
                         "algorithmic_gflop_per_step": pf.value / args.steps / 1e9,
                         "algorithmic_gflop_per_step_expected": conv_flop_algorithmic / 1e9 if conv_flop_algorithmic else None,
                         "algorithmic_bytes_per_launch": pb.value / pn.value if pn.value else None,
                         "executed_mma_factor": ("3 on the 36 forward clips, 1 on the 24 data-gradient clips"
                                                 if trainer.plan.precision == "tf32x3" else 1),
                         "peak_source": peaks["source"] + " bf16 sustained (tf32 nominal dense peak is half of bf16)",
                         "launches": pn.value, "launches_per_step": pn.value / args.steps,
                         "ms_per_step": pm.value / args.steps,
                         "share_of_step": pm.value / ms_eager if ms_eager else None,
                         "step_tflops": step_flop * args.steps / (ms * 1e-3) / 1e12},
            "frontend": {"metric": "log-mel frontend", "clips_per_s": fe_cps, "algorithmic_GBps": fe_cps * CLIP_BYTES_FRONTEND / 1e9,
                         "hbm_frac": fe_cps * CLIP_BYTES_FRONTEND / 1e9 / peaks["hbm"] / world,
                         "stft_mel_kernel_ms_per_256_clips": fm.value / max(1, fn_.value),
                         "stft_mel_fp32_tflops": 256 * 1255 * 70000.0 / (fm.value / max(1, fn_.value) * 1e-3) / 1e12,
                         # the roof that applies: 70 kFLOP of fp32 work per frame (2048-point real FFT, untangle, magnitude,
                         # interval projection) = 88 MFLOP per clip; at 70 % of the HBM copy rate the kernel would have to
                         # sustain 2.4 M clips/s = 209 TFLOP/s, 2.8x the fp32 FMA peak below
                         "fp32_fma_peak_tflops": 148 * 128 * 2 * 1.965e9 / 1e12,
                         "stft_mel_frac_of_fp32_peak": 256 * 1255 * 70000.0 / (fm.value / max(1, fn_.value) * 1e-3) / 1e12
                                                        / (148 * 128 * 2 * 1.965e9 / 1e12),
                         "bound": "fp32 / shared-memory pipes (ncu: issue slots 50-54 %, LSU shared-memory pipe 71-75 % busy, "
                                  "profiles/r01g_frontend_kernels.md); HBM roofline at 70 % would need 209 TFLOP/s fp32",
                         "db_transform": {"bound": "hbm", "achieved_GBps": db_gbps, "peak_GBps": peaks["hbm"],
                                          "frac": db_gbps / peaks["hbm"] if db_gbps else None,
                                          "ms_per_256_clips": dm.value / max(1, dn.value),
                                          "note": "algorithmic bytes (amplitude-mel read once + log-mel written once); the per-clip "
                                                  "80 dB clamp needs the clip maximum first, so the kernel pair reads the mel twice"}},
            "single_pass_tf32" if other == "tf32" else "parity_mode_tf32x3": other_mode,
            "parity": parity,
            "gpu_eager_baseline": eager,
            "cpu_baseline": cpu_baseline,
            "clocks": clocks,
        }
        out.emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_ada(args, out):
    """BASELINE.json configs[2]: the SCMT + adversarial-domain-adaptation iteration (src/main_scmt_ada_weak_seperate.py:
    train_mt with a discriminator) -- adversarial update (student forward of both domains -> gradient reversal ->
    Clip_Discriminator -> BCE -> optimizer_crnn / optimizer_d), then the mean-teacher update with the weak labels of the
    real batch; SGD-Nesterov x 3; per GPU 12 synthetic + 12 real (6 weak-labelled + 6 pseudo-labelled) clips."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from bsed_b200 import _lib, engine
    from bsed_b200.DA.cdan_frame import ConditionalDomainAdversarialLoss
    from bsed_b200.main import AdaptationTrainer
    from bsed_b200.models import CRNN, Predictor
    from bsed_b200.models.CRNN import Clip_Discriminator
    from bsed_b200.utilities import synth
    from bsed_b200.utilities.utils import weights_init
    lib = _lib.load()
    dev = torch.device("cuda", local)
    torch.manual_seed(2023 + rank)

    def make():
        m, p = CRNN(**engine.REFERENCE_CRNN_KWARGS), Predictor(**engine.REFERENCE_PREDICTOR_KWARGS)
        weights_init(m)
        weights_init(p)
        return m.to(dev).train(), p.to(dev).train()

    model, predictor = make()
    ema_model, ema_predictor = make()
    for prm in list(ema_model.parameters()) + list(ema_predictor.parameters()):
        prm.detach_()
    torch.manual_seed(7)                                   # identical discriminator on every rank
    disc = Clip_Discriminator(256).to(dev).train()
    crit = ConditionalDomainAdversarialLoss(disc)
    tr = AdaptationTrainer(model, predictor, ema_model, ema_predictor, crit, lr=5e-4, momentum=0.9, weight_decay=1e-4,
                           n_syn=N_SYN, n_real=N_REAL, dropout_seed=2023 + rank)
    xs = torch.from_numpy(synth.make_logmel_like(N_SYN, seed=3 + rank)).to(dev)
    x = torch.from_numpy(synth.make_logmel_like(N_REAL, seed=40 + rank)).to(dev)
    x_ema = (x + 0.05 * torch.from_numpy(synth.make_logmel_like(N_REAL, seed=60 + rank)).to(dev)).contiguous()
    ts = torch.from_numpy(synth.make_targets(N_SYN, seed=5 + rank)).to(dev)
    tw = torch.from_numpy(synth.make_targets(N_REAL, seed=8 + rank)).max(1)[0].to(dev)
    ramp = 50 * 100

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, warmup, before=None):
        for i in range(warmup):
            fn(i)
        sync_all()
        if before is not None:
            before()
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
    cnt = {}
    ms = timed(lambda i: tr.step(x, x_ema, tw, xs, ts, i, ramp), args.steps, args.warmup,
               lambda: cnt.update(l0=lib.bsed_launch_count()))
    launches = lib.bsed_launch_count() - cnt["l0"]
    value = (N_SYN + N_REAL) * world * args.steps / (ms * 1e-3)
    # end to end: inputs from pinned host memory every step, losses read back
    host = [t.cpu().pin_memory() for t in (x, x_ema, tw, xs, ts)]
    dbuf = [torch.empty_like(t) for t in (x, x_ema, tw, xs, ts)]
    hloss = torch.empty(5).pin_memory()

    def e2e_step(i):
        for d_, h_ in zip(dbuf, host):
            d_.copy_(h_, non_blocking=True)
        losses, dom = tr.step(dbuf[0], dbuf[1], dbuf[2], dbuf[3], dbuf[4], 1000 + i, ramp)
        hloss[:4].copy_(losses, non_blocking=True)
        hloss[4:].copy_(dom, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    ms_e2e = timed(e2e_step, args.steps, 2)
    h2d_bytes = sum(t.numel() * t.element_size() for t in host)
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)
        eager = cpu = None
        if world == 1:
            eager = _eager_ada(dev, 10, 3)
            cpu = _eager_ada(torch.device("cpu"), 2, 1)
        out.emit(json.dumps({
            "metric": "SCMT + ADA CRNN train clips/s", "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": tr.plan.precision, "data": "synthetic",
            "config": {"workload": "SCMT + adversarial domain adaptation step (src/main_scmt_ada_weak_seperate.py:train_mt with "
                                   "Clip_Discriminator + cdan_frame + gradient reversal): per GPU 12 synthetic + 12 real clips; "
                                   "adversarial update (2 student calls fwd+bwd, D fwd+bwd, 2 x SGD-Nesterov) then the mean-teacher "
                                   "update (3 calls, BCE strong/weak incl. the real batch's weak labels, MSE consistency, "
                                   "SGD-Nesterov, state-dict EMA)", "model": "crnn + clip_discriminator", "clips_per_step_per_gpu": 24,
                       "parallelism": f"dp{world}: " + ("three fused peer-memory reduce + update kernels per step (encoder+predictor, "
                                                        "encoder-adversarial, discriminator)" if tr.dp is not None else
                                                        "NCCL all-reduce" if world > 1 else "single GPU"),
                       "discriminator_precision": disc.precision},
            "e2e": {"value": (N_SYN + N_REAL) * world * args.steps / (ms_e2e * 1e-3), "unit": "clips/s",
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 20, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "gpu_eager_baseline": eager, "cpu_baseline": cpu, "clocks": sampler.summary()}))
    if world > 1:
        dist.destroy_process_group()


def _eager_ada(dev, steps, warmup):
    """The reference's module structure (oracle restatement, stock nn.Dropout) running oracle/train.py:ada_step in PyTorch
    eager mode on `dev` (the B200 with stock cuDNN / cuBLAS settings, or the host cores)."""
    import torch
    from oracle import da as oda
    from oracle import train as otrain
    from bsed_b200.utilities import synth
    if dev.type == "cpu":
        torch.set_num_threads(os.cpu_count() or 1)
    oc, op, tc, tp = _oracle_models(False)
    for m in (oc, tc):
        _stock_dropout(m)
    od = oda.OracleClipDiscriminator()
    oda.seeded_disc_init(od, 3)
    for m in (oc, op, tc, tp, od):
        m.to(dev).train()
    xs, xr, xr_ema, ts = [t.to(dev) for t in _step_inputs(N_SYN, N_REAL)]
    tw = torch.from_numpy(synth.make_targets(N_REAL, seed=8)).max(1)[0].to(dev)
    sgd = dict(lr=5e-4, momentum=0.9, weight_decay=1e-4, nesterov=True)
    opt = torch.optim.SGD(list(oc.parameters()) + list(op.parameters()), **sgd)
    opt_c, opt_d = torch.optim.SGD(oc.parameters(), **sgd), torch.optim.SGD(od.parameters(), **sgd)

    def step(i):
        loss, _, _ = otrain.ada_step(oc, op, tc, tp, od, opt, opt_c, opt_d, xr, xr_ema, tw, xs, ts, i, 5000, i)
        return loss.item()

    for i in range(warmup):
        step(i)
    if dev.type == "cuda":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        step(warmup + i)
    if dev.type == "cuda":
        torch.cuda.synchronize()
    sec = (time.perf_counter() - t0) / steps
    r = dict(value=(N_SYN + N_REAL) / sec, unit="clips/s", ms_per_step=sec * 1e3, steps=steps, warmup=warmup,
             kind="port (oracle/train.py:ada_step over oracle/crnn.py + oracle/da.py = the reference's torch.nn layers)")
    if dev.type == "cpu":
        r.update(cores=torch.get_num_threads(), sample=f"{steps} steps of the full 12 + 12 (+ 12 teacher) clip iteration")
    else:
        r.update(cudnn_allow_tf32=bool(torch.backends.cudnn.allow_tf32), matmul_allow_tf32=bool(torch.backends.cuda.matmul.allow_tf32))
    return r


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
    from bsed_b200.utilities import shard
    base = synth.make_clips(24, seed=11)                                   # the stream is these 24 clips over and over
    n_clips = 360 * max(1, args.replicate)                                 # 360 clips = 1 h at 32 kHz
    begin, end = shard.clip_shard(n_clips, rank, world)
    # every rank holds ITS contiguous block of the stream in pinned host memory (a x64 stream is 29 GB: only the shard
    # is materialised); pseudo_label_stream then walks that block exactly as it walks its shard of a whole stream
    idx = np.arange(begin, end) % 24
    audio = torch.from_numpy(np.ascontiguousarray(base[idx]).reshape(-1)).pin_memory()
    run = lambda: pseudo_label_stream(audio, m, p, batch_clips=48, rank=0, world=1, gather=False,
                                      name_fmt="stream_{:05d}" if begin == 0 else "stream_" + str(begin) + "+{:05d}")
    for _ in range(max(1, args.warmup)):
        run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from bsed_b200 import _lib
    lib = _lib.load()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lib.bsed_launch_count()
    e0.record()
    reps = max(1, args.steps // 4)
    for _ in range(reps):
        res = run()
    e1.record()
    torch.cuda.synchronize()
    launches = lib.bsed_launch_count() - l0
    sampler.stop_flag = True
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    cpu = None
    if rank == 0 and world == 1:
        # the reference's arithmetic for the same pipeline on the host cores (test-infrastructure leg): oracle frontend clip by
        # clip, the torch.nn CRNN + Predictor in eval mode, threshold -> median filter -> events
        import time as _time
        from oracle import crnn as ocrnn, frontend as ofe, postproc as opp
        oc = ocrnn.OracleCRNN(**{**ocrnn.CRNN_KWARGS, "dropout": 0.5}).eval()
        op = ocrnn.OraclePredictor(**ocrnn.PREDICTOR_KWARGS).eval()
        ocrnn.reference_style_init(oc, op, 1, 0.2)
        n_cpu = 48
        t0 = _time.perf_counter()
        with torch.no_grad():
            xc = np.stack([ofe.logmel(base[i % 24]) for i in range(n_cpu)])[:, None]
            s_or, w_or = op(oc(torch.from_numpy(xc))[0])
            n_ev = sum(len(opp.events_from_strong(s_or[i].numpy(), 0.5, 14)) for i in range(n_cpu))
        dt = _time.perf_counter() - t0
        cpu = {"value": n_cpu / dt, "unit": "clips/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": "%d clips: oracle/frontend.py log-mel, oracle/crnn.py torch.nn eval forward, oracle/postproc.py events "
                         "(%.1f s of CPU work, %d events)" % (n_cpu, dt, n_ev)}
    if rank == 0:
        cps = n_clips * reps / (float(ms) * 1e-3)
        shard_clips = end - begin
        out.emit(json.dumps({"metric": "log-mel + CRNN pseudo-label inference clips/s", "value": cps, "unit": "clips/s",
                          "n_gpus": world, "steps": reps, "warmup": args.warmup, "ms_per_step": float(ms) / reps,
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                          "dtype": m.precision or engine.default_precision(), "data": "synthetic",
                          "config": {"workload": "%d x 10 s clips of synthetic audio (%.1f h) from pinned host memory -> framed STFT "
                                                 "-> mel -> dB -> CRNN + Predictor eval -> weak labels + median-filtered events "
                                                 "(pseudo_labeling.pseudo_label_stream), clips sharded over ranks" % (n_clips, n_clips / 360),
                                     "events_rank0": len(res["events"]), "audio_GBps": cps * 1280000 / 1e9,
                                     "timing": "inputs (46 MB per rank and pass at x1) stream from pinned host memory inside the "
                                               "timed region; stream larger than L2"},
                          # the timed call IS the public API with host buffers: audio H2D and label / event D2H every pass
                          "e2e": {"value": cps, "unit": "clips/s", "h2d_bytes_per_step": shard_clips * 1280000,
                                  "d2h_bytes_per_step": shard_clips * (20 * ((1255 // 4 + 1) // 2) * 3 * 4 + 4 + 20)},
                          "gpu_launches": int(launches), "cpu_baseline": cpu, "clocks": sampler.summary()}))
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
    ap.add_argument("--workload", default="train", choices=["train", "pseudo_label", "ada"],
                    help="train = the headline mean-teacher step (default); pseudo_label = configs[4] inference pipeline; "
                         "ada = configs[2], the SCMT + adversarial domain adaptation iteration")
    ap.add_argument("--replicate", type=int, default=1, help="pseudo_label: repeat the 1 h stream this many times")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return
    with StdoutGuard() as out:
        if args.workload == "pseudo_label":
            run_pseudo_label(args, out)
        elif args.workload == "ada":
            run_ada(args, out)
        else:
            run_b200(args, out)


if __name__ == "__main__":
    main()
