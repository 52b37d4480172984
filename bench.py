#!/usr/bin/env python
"""Benchmark of the ELBO train step (BASELINE.json metric: captions/sec).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference ...                      # CPU restatement of the reference graph

Workload (config.workload): BASELINE.json configs[1] -- Normal-prior CVAE, bf16 operands, on-device
VGG16 forward, batch = 256 images x 5 captions = 1280 captions per step per GPU, T = 20, V = 11313.
Under torchrun every rank runs the same per-GPU batch (weak scaling) and the gradients are summed
with one all-reduce; the printed value is the whole-job captions/s.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (images per step per GPU, on-device VGG forward, prior, c_v)
    "cfg2_vgg_normal_b256": dict(B=256, vgg=True, prior="Normal", c_v=False),
    "cfg1_feats_normal_b32": dict(B=32, vgg=False, prior="Normal", c_v=False),
    "feats_normal_b256": dict(B=256, vgg=False, prior="Normal", c_v=False),
    "cfg3_feats_gmm_cv_b128": dict(B=128, vgg=False, prior="GMM", c_v=True),
    "cfg4_finetune_ag_cv_b256": dict(B=256, vgg=True, prior="AG", c_v=True, fine_tune=True),
    "finetune_ag_cv_b64": dict(B=64, vgg=True, prior="AG", c_v=True, fine_tune=True),
    # BASELINE configs[4]: inference path -- greedy + beam-5 decode over 40k synthetic val images (one step = both
    # decodes of one batch of B feature rows; metric images/s)
    "cfg5_decode_greedy_beam5": dict(B=1024, vgg=False, prior="Normal", c_v=False, decode=True, images=40000),
}
DEFAULT_WORKLOAD = "cfg2_vgg_normal_b256"
C, T, V = 5, 20, 11313


def params_for(w):
    from vae_captioning_b200.parameters import Parameters
    p = Parameters()
    p.prior = w["prior"]
    p.use_c_v = w["c_v"]
    p.batch_size = w["B"]
    p.vocab_size = V
    p.fine_tune = bool(w.get("fine_tune"))
    return p


# ------------------------------------------------------------------------------------------------
# algorithmic work per kernel family per step: FLOPs for tensor-bound, bytes for HBM-bound
VGG = [(224, 3, 64), (224, 64, 64), (112, 64, 128), (112, 128, 128), (56, 128, 256), (56, 256, 256), (56, 256, 256),
       (28, 256, 512), (28, 512, 512), (28, 512, 512), (14, 512, 512), (14, 512, 512), (14, 512, 512)]


def family_work(p, B, has_enc=True, n_params=None):
    N, E, He, Hd, Z, S, F = B * C, p.embed_size, p.encoder_hidden, p.decoder_hidden, p.latent_size, p.gen_z_samples, 4096
    cv = 1 if p.use_c_v else 0
    enc_steps, dec_steps = T + 1 + cv, T + 2 + cv
    n_adam = (F * E + E + (E + He) * 4 * He + 4 * He + 2 * (He * Z + Z) + (E + Hd) * 4 * Hd + 4 * Hd + Z * S * E + E +
              Hd * V + V + 2 * V * E)
    if n_params is not None:  # every variable an optimiser updates (GMM / AG heads, cv_emb, the cnn/ scope when fine-tuning)
        n_adam = n_params
    w = {
        "conv": ("tensor", sum(2.0 * hw * hw * 9 * ci * co for hw, ci, co in VGG) * B),
        "fc": ("tensor", 2.0 * B * (25088 * 4096 + 4096 * 4096)),
        "fc_dgrad": ("tensor", 2.0 * B * (25088 * 4096 + 4096 * 4096)),
        "fc_wgrad": ("tensor", 2.0 * B * (25088 * 4096 + 4096 * 4096)),
        "conv_wgrad": ("tensor", sum(2.0 * hw * hw * 9 * ci * co for hw, ci, co in VGG) * B),
        "conv_dgrad": ("tensor", sum(2.0 * hw * hw * 9 * ci * co for hw, ci, co in VGG[1:]) * B),
        "lstm_fwd_step": ("tensor", 2.0 * N * ((E + He) * 4 * He * enc_steps + (E + Hd) * 4 * Hd * dec_steps)),
        "lstm_bwd_step": ("tensor", 2.0 * N * (He * 4 * He * (enc_steps - 1) + Hd * 4 * Hd * (dec_steps - 1))),
        "lstm_wgrad": ("tensor", 2.0 * N * ((E + He) * 4 * He * enc_steps + (E + Hd) * 4 * Hd * dec_steps)),
        "lstm_dx": ("tensor", 2.0 * N * (E * 4 * He * enc_steps + E * 4 * Hd * dec_steps)),
        "logits_fwd": ("tensor", 2.0 * N * T * Hd * V),
        "logits_dgrad": ("tensor", 2.0 * N * T * Hd * V),
        "logits_wgrad": ("tensor", 2.0 * N * T * Hd * V),
        "z_rnn": ("tensor", 2.0 * N * S * Z * E),
        "z_rnn_dgrad": ("tensor", 2.0 * N * S * Z * E),
        "z_rnn_wgrad": ("tensor", 2.0 * N * S * Z * E),
        "imf_emb": ("tensor", 2.0 * B * F * E),
        "imf_emb_bwd": ("tensor", 2.0 * B * F * E),
        # HBM-bound stages: bytes that must move
        "ce": ("hbm", 2.0 * N * T * V * 2),            # bf16 logits read once, dlogits written once
        "adam": ("hbm", 28.0 * n_adam),                 # read p,g,m,v; write p,m,v (fp32)
        "sumsq": ("hbm", 4.0 * n_adam),
        "sample_z": ("hbm", S * N * Z * 2.0 + 2 * N * Z * 4.0),
        "dz_reduce": ("hbm", S * N * Z * 4.0),
    }
    # the whole-sequence LSTM kernels run every step's 4-gate GEMM in one launch: same algorithmic FLOPs
    w["lstm_fwd_seq"] = w["lstm_fwd_step"]
    w["lstm_bwd_seq"] = w["lstm_bwd_step"]
    names = ["conv1_1", "conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv3_3", "conv4_1", "conv4_2", "conv4_3",
             "conv5_1", "conv5_2", "conv5_3"]
    for nm, (hw, ci, co) in zip(names, VGG):
        w[nm] = ("tensor", 2.0 * hw * hw * 9 * ci * co * B)
    return w


def total_flops(p, B, vgg):
    w = family_work(p, B)
    cvae = ("lstm_fwd_step", "lstm_bwd_step", "lstm_wgrad", "lstm_dx", "logits_fwd", "logits_dgrad", "logits_wgrad", "z_rnn",
            "z_rnn_dgrad", "z_rnn_wgrad", "imf_emb", "imf_emb_bwd")
    tot = sum(w[k][1] for k in cvae)
    heads = 1 if p.prior == "Normal" else 3  # minimum-algorithmic (active heads only)
    tot += 3 * 2.0 * B * C * p.encoder_hidden * 2 * p.latent_size * heads
    if vgg:
        tot += (w["conv"][1] + w["fc"][1]) * (3 if p.fine_tune else 1)  # fine-tune: + dgrad + wgrad
    return tot


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self._stop_evt = threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------
def run_reference(args, w, name):
    """The reference's CPU path: the oracle (torch-CPU fp32 restatement; TF1/zhusuan cannot be installed)
    on a bounded sample of the workload, all host threads."""
    import torch
    from oracle import cvae_oracle as O
    cores = cpu_cores()
    torch.set_num_threads(cores)
    Bs = args.ref_batch
    cfg = O.Config(prior=w["prior"], use_c_v=w["c_v"], vocab_size=V, fine_tune=bool(w.get("fine_tune")))
    params = O.init_params(cfg, seed=1, with_cnn=w["vgg"], dtype=torch.float32)
    batch = O.synthetic_batch(cfg, Bs, T, seed=0, dtype=torch.float32, with_images=w["vgg"])
    opt = {"t": 0, "m": {}, "v": {}}

    def step():
        b = dict(batch)
        if w["vgg"] and not cfg.fine_tune:
            with torch.no_grad():
                b["feats"] = O.vgg16_fc2(params, batch["images"])
        O.train_step(params, opt, cfg, b)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = Bs * C / dt
    sample = "%d images x %d captions per step (full workload: %d images), fp32 torch-CPU restatement of the TF1 graph" % (
        Bs, C, w["B"])
    return {"metric": "captions/sec (ELBO train step)", "value": val, "unit": "captions/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": name, "images_per_step": Bs, "captions_per_step": Bs * C, "seq_len": T, "vocab": V},
            "cpu_baseline": {"value": val, "unit": "captions/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "captions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


# ------------------------------------------------------------------------------------------------
class _SynthDict(object):
    """Stand-in for utils/captions.py Dictionary: ids only (<PAD>=0, <BOS>=1, <EOS>=2)."""

    def __init__(self, vocab):
        self.idx2word = {i: "w%d" % i for i in range(vocab)}
        self.idx2word.update({0: "<PAD>", 1: "<BOS>", 2: "<EOS>"})
        self.word2idx = {w: i for i, w in self.idx2word.items()}
        self.vocab_size = vocab


def run_decode_reference(args, w, name):
    """CPU restatement of the reference generation loops (one sess.run per token per beam at batch 1,
    vae_model/decoder.py:145-320) on a bounded sample of images, numpy float64 on the host cores."""
    import torch
    from oracle import cvae_oracle as O
    from oracle import decode_oracle as D
    cores = cpu_cores()
    torch.set_num_threads(cores)
    cfg = O.Config(vocab_size=V)
    params = O.init_params(cfg, seed=1)
    model = D.GenModel(params, cfg)
    g = np.random.Generator(np.random.PCG64(0))
    n_img = max(1, args.ref_batch // 8)
    feats = np.maximum(0, g.standard_normal((n_img, 4096))).astype(np.float32)

    def step():
        for i in range(n_img):
            eps = g.standard_normal((cfg.gen_z_samples, cfg.latent_size))
            D.online_inference(model.make_step(feats[i], None, eps), "greedy", cfg.gen_max_len, 1.0, 1, 2)
            D.beam_search(model.make_step(feats[i], None, eps), 5, cfg.gen_max_len, 1, 2)

    for _ in range(min(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = n_img / dt
    sample = "%d images per step, greedy + beam-5 each, numpy restatement of the per-token sess.run loops" % n_img
    return {"metric": "images/sec (greedy + beam-5 decode)", "value": val, "unit": "images/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp64", "data": "synthetic",
            "config": {"workload": name, "images_per_step": n_img, "gen_max_len": 30, "beam": 5, "vocab": V},
            "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def run_decode(args, w, name, rank, world, local_rank):
    """cfg 5: Decoder.online_inference('greedy') + Decoder.beam_search(beam 5) (vae_model/decoder.py:145-320) batched on
    the device. Replicas only: every rank decodes its own images, no collective."""
    import torch
    import torch.distributed as dist
    from vae_captioning_b200 import lib as L
    from vae_captioning_b200.engine import Engine
    from vae_captioning_b200.decode import Decoder
    from vae_captioning_b200 import synthetic
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = L.load()
    lib.vc_launch_count.restype = ctypes.c_ulonglong
    p = params_for(w)
    p.mode = "inference"
    B = w["B"]
    eng = Engine(p, vocab_size=V, max_batch=B, max_len=T, device=local_rank)
    eng.load_state(synthetic.init_weights(eng.variables(), seed=1))
    dec = Decoder(eng, p, _SynthDict(V))
    g = np.random.Generator(np.random.PCG64(rank))
    feats = np.maximum(0, g.standard_normal((B, 4096), dtype=np.float32))
    host = torch.from_numpy(feats).pin_memory()

    def step():
        dec.greedy_tokens(host.numpy(), rng={"seed": 7})
        dec.beam_tokens(host.numpy(), beam_size=5, rng={"seed": 7})

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = lib.vc_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = int(lib.vc_launch_count() - l0)
    clocks = sampler.stop() if sampler else None
    value = world * B * args.steps / (ms / 1e3)
    # roofline of the dominant kernel family: the per-token vocabulary projection [M, 512] x [512, V] -> fp32 logits. Its
    # algorithmic bytes per launch: W_o once (bf16) + the logits rows it must hand to the top-k (fp32) + the h rows (bf16);
    # M = B rows for greedy, up to B * beam for the beam steps (summed over the launches of one profiled step)
    roofline, families = None, {}
    if rank == 0 and not args.no_profile:
        torch.cuda.synchronize()
        lib.vc_profile_enable(1)
        step()
        names = ctypes.create_string_buffer(8192)
        msb = (ctypes.c_float * 256)()
        cnt = (ctypes.c_int * 256)()
        n = lib.vc_profile_collect(names, 8192, msb, cnt, 256)
        lib.vc_profile_enable(0)
        for i, fn in enumerate(names.value.decode().split(",") if n else []):
            families[fn] = {"ms_per_step": msb[i], "launches_per_step": cnt[i]}
        if "logits_decode" in families:
            f = families["logits_decode"]
            H = p.decoder_hidden
            n_g, n_b = 30, 30  # launches: greedy gen_max_len steps with M = B, beam steps with M = B * beam (first beam step M = B)
            rows = n_g * B + B + (n_b - 1) * B * 5
            nbytes = f["launches_per_step"] * V * H * 2.0 + rows * (V * 4.0 + H * 2.0)
            peak = 6454.6
            try:
                peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", peak)
            except Exception:
                pass
            ach = nbytes / (f["ms_per_step"] / 1e3) / 1e9
            roofline = {"kernel": "logits_decode", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                        "frac": ach / peak, "traffic": None, "share_of_step": f["ms_per_step"] / sum(x["ms_per_step"] for x in families.values()),
                        "note": "algorithmic bytes = W_o (bf16) per launch + fp32 logits rows written for the top-k + bf16 state rows read"}
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ra = argparse.Namespace(**vars(args))
        ra.steps, ra.warmup = 1, 0
        cpu_baseline = run_decode_reference(ra, w, name)["cpu_baseline"]
    if rank == 0:
        # weight-streaming view of the dominant kernel (the [M, 512] x [512, V] vocabulary projection per token):
        # algorithmic bytes per decoded token row = its logits row (bf16) + the row's share of W_o
        line = {"metric": "images/sec (greedy + beam-5 decode)", "value": value, "unit": "images/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": name, "images_per_step_per_gpu": B, "gen_max_len": 30, "beam": 5, "vocab": V,
                           "val_images": w["images"], "seconds_for_val_set": w["images"] / value,
                           "note": "untrained seeded weights: most captions run to gen_max_len (worst case)",
                           "parallelism": "replicas%d" % world},
                "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 2 * int(host.numel()) * 4,
                        "d2h_bytes_per_step": B * 30 * 4 + B * 5 * 30 * 4 + B * 5 * 8 + B * 8},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
                "families": families}
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    eng.close()
    return 0


_REAL_STDOUT = None


def quiet_stdout():
    """Library chatter on stdout (NCCL's version banner, torchrun notices) would precede the result line: send fd 1 to
    stderr for the duration of the run and keep the real stdout for the ONE JSON line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def _crc_state(eng):
    """crc32 over every variable of the handle (host copy): equal on all ranks iff the replicas stayed in lock step."""
    import zlib
    crc = 0
    for name, _, _ in eng.variables():
        crc = zlib.crc32(eng.get_variable(name).tobytes(), crc)
    return crc


def run_train(args, name, rank, world, local_rank, steps, warmup, full):
    """One workload through the CUDA path: device-resident leg (`value`), host-buffer leg (`e2e`), at world > 1 the
    all-reduce accounting and the replica check; with `full` also the per-kernel-family roofline legs and the CPU
    baseline. Returns the JSON line (rank 0) or None."""
    import torch
    import torch.distributed as dist
    from vae_captioning_b200 import lib as L
    from vae_captioning_b200.engine import Engine
    from vae_captioning_b200 import synthetic

    w = WORKLOADS[name]
    lib = L.load()
    lib.vc_launch_count.restype = ctypes.c_ulonglong
    p = params_for(w)
    B = w["B"]
    N = B * C
    eng = Engine(p, vocab_size=V, max_batch=B, max_len=T, device=local_rank, with_cnn=w["vgg"])
    eng.load_state(synthetic.init_weights(eng.variables(), seed=1))
    n_params = sum(int(np.prod(sh)) for _, sh, tr in eng.variables() if tr)
    if w["prior"] == "AG":  # init_clusters (utils/vae_utils.py:6-31): seeded stand-in for ./pickles/cluster_means.pickle
        eng.set_cluster_means(np.random.Generator(np.random.PCG64(2)).standard_normal((90, p.latent_size)).astype(np.float32))
    if world > 1:  # the handle's own NCCL communicator: bucketed all-reduce inside every train step (include/vaecap.h)
        eng.attach_comm(rank, world)
    feed = synthetic.make_batch(B, C, T, V, seed=rank, images=w["vgg"], cluster_vectors=w["c_v"] or w["prior"] != "Normal")

    # pinned host buffers (e2e leg) and device-resident copies (value leg)
    host = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in feed.items()}
    dev = {k: v.cuda(non_blocking=True) for k, v in host.items()}
    if w["vgg"]:  # the e2e leg feeds the pixels as uint8, the dtype of the reference's HDF5 image store (batch_gen.py:278-294)
        host["image_f_inputs"] = torch.from_numpy(feed["image_f_inputs"].astype(np.uint8)).pin_memory()
    torch.cuda.synchronize()

    step_no = [0]
    e2e_primed = []
    e2e_pending = [False]
    tail_out = []

    # Frozen extractor (cfg 2): the VGG16 forward of batch i+1 does not depend on step i (its weights never change), so it
    # runs on a second stream while the caption model of batch i runs on the main one -- the same schedule
    # vc_stage_batch / vc_train_step_staged give the host-buffer leg. Two fc2 buffers, events both ways.
    pipelined = bool(w["vgg"] and not w.get("fine_tune"))
    side = torch.cuda.Stream() if pipelined else None
    fc2_buf = [torch.empty((B, 4096), dtype=torch.float32, device="cuda") for _ in range(2)] if pipelined else None
    fc2_ready = [torch.cuda.Event() for _ in range(2)] if pipelined else None
    fc2_free = [torch.cuda.Event() for _ in range(2)] if pipelined else None
    fc2_primed = []

    def vgg_ahead(i):
        with torch.cuda.stream(side):
            if len(fc2_primed) >= 2:
                side.wait_event(fc2_free[i & 1])  # the step that read this buffer two steps ago
            eng.vgg_forward_device(dev["image_f_inputs"], out=fc2_buf[i & 1])
            fc2_ready[i & 1].record(side)
        fc2_primed.append(i)

    def step_device(serial=False):
        feats = dev["image_f_inputs"]
        if pipelined and serial:  # per-kernel timing leg: one stream, kernels timed alone
            feats = eng.vgg_forward_device(feats)
        elif pipelined:
            i = step_no[0]
            if not fc2_primed:
                vgg_ahead(i)
            main = torch.cuda.current_stream()
            main.wait_event(fc2_ready[i & 1])
            feats = fc2_buf[i & 1]
            vgg_ahead(i + 1)
        # at world > 1 this same call all-reduces the gradient buckets under the backward pass and applies the mean
        eng.train_step_device(feats, dev["ann_inputs_enc"], dev["ann_inputs_dec"], dev["ann_lengths"], step_no[0],
                              c_i=dev.get("c_i"), rng={"seed": 1234 + rank}, fetch=False)
        if pipelined and not serial:
            fc2_free[step_no[0] & 1].record(torch.cuda.current_stream())
        step_no[0] += 1

    def step_e2e():
        # public API with HOST buffers: H2D of the step's inputs and D2H of the loss inside the timed region.
        # Double-buffered feed (vc_stage_batch / vc_train_step_staged): every call copies ONE batch from pinned host
        # memory (the next step's, on the copy stream, overlapping this step's compute) and runs ONE step whose
        # scalars are read back; the first batch is staged by the untimed warm-up calls
        i = step_no[0]
        stage = lambda slot: eng.stage_batch(slot, host["image_f_inputs"].numpy(), host["ann_inputs_enc"].numpy(),
                                             host["ann_inputs_dec"].numpy(), host["ann_lengths"].numpy(),
                                             c_i=host["c_i"].numpy() if "c_i" in host else None,
                                             images=w["vgg"] and not w.get("fine_tune"))
        if not e2e_primed:
            stage(i & 1)
            e2e_primed.append(True)
        stage((i + 1) & 1)
        # the step's scalars are read back EVERY step, one step late (vc_step_result_queue / vc_step_result): the host
        # enqueues step i + 1 before it waits for the loss of step i, so its launch time hides behind the GPU's work
        eng.train_step_staged(i & 1, i, rng={"seed": 1234 + rank}, fetch=False)
        eng.queue_result()
        step_no[0] += 1
        out = eng.pop_result() if e2e_pending[0] else None
        e2e_pending[0] = True
        return out

    def e2e_drain():
        # the last step's result, still inside the timed region
        if e2e_pending[0]:
            e2e_pending[0] = False
            return eng.pop_result()
        return None

    def timed(fn, k, tail=None):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        if tail is not None:
            tail_out.append(tail())
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)  # the look-ahead forward of the last step is inside the region
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.barrier()
            ms = float(t.item())
        return ms

    for _ in range(max(warmup, 3)):
        step_device()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = lib.vc_launch_count()
    ms = timed(step_device, steps)
    launches = int(lib.vc_launch_count() - l0)
    clocks = sampler.stop() if sampler else None
    value = world * N * steps / (ms / 1e3)
    comm = eng.comm_stats() if world > 1 else None
    ms_serial = None
    if pipelined and full:  # the same K steps on one stream, for the record (explains what the overlap buys)
        torch.cuda.synchronize()
        fc2_primed.clear()
        ms_serial = timed(lambda: step_device(serial=True), steps)

    # e2e leg
    last, ms_e2e, e2e_val = None, float("nan"), None
    if not args.no_e2e:
        for _ in range(2):
            step_e2e()
        e2e_drain()
        ms_e2e = timed(step_e2e, steps, tail=e2e_drain)
        last = tail_out[-1]
        e2e_val = world * N * steps / (ms_e2e / 1e3)
    h2d = sum(v.numel() * v.element_size() for v in host.values())

    # data-parallel correctness and what the collective costs (world > 1)
    dp_check, allreduce, loss_mean = None, None, None
    if world > 1:
        torch.cuda.synchronize()
        crcs = [None] * world
        dist.all_gather_object(crcs, _crc_state(eng))
        dp_check = "ok" if len(set(crcs)) == 1 else "MISMATCH %s" % (crcs,)
        if last is not None:  # every rank reports its tower's loss; the job's loss is their mean
            losses = [None] * world
            dist.all_gather_object(losses, float(last["rec_loss"]))
            loss_mean = float(np.mean(losses))
        k2 = max(3, min(steps, 10))
        fc2_primed.clear()
        eng.comm_set_mode(2)  # one all-reduce of the whole buffer behind the backward pass (round 1's schedule)
        for _ in range(2):
            step_device()
        ms_m2 = timed(step_device, k2) / k2
        eng.comm_set_mode(0)  # no reduction at all: the ranks diverge from here on, timing only (last leg)
        for _ in range(2):
            step_device()
        ms_m0 = timed(step_device, k2) / k2
        eng.comm_set_mode(1)
        allreduce = {"bytes": comm["bytes"], "buckets": comm["buckets"], "span_ms": comm["span_ms"],
                     "exposed_ms": ms / steps - ms_m0, "exposed_ms_unbucketed": ms_m2 - ms_m0,
                     "ms_per_step_no_allreduce": ms_m0, "ms_per_step_unbucketed": ms_m2,
                     "transport": "bf16 (VC_GRAD_BF16=1)" if os.environ.get("VC_GRAD_BF16") == "1" else "fp32",
                     "note": "span = first bucket start to last bucket end on the communication stream of the last timed "
                             "step; exposed = step time minus the same step with the reduction switched off"}

    # per-kernel-family timing with CUDA events on the launching stream (extra steps after the timed region)
    roofline = None
    lstm_roofline = None
    families = {}
    if full and rank == 0 and world == 1 and not args.no_profile:  # extra steps on one rank only would strand its all-reduce
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        psteps = 3
        torch.cuda.synchronize()
        fc2_primed.clear()
        lib.vc_profile_enable(1)
        for _ in range(psteps):
            step_device(serial=True)  # un-pipelined, so every kernel's events bracket it running alone
        names = ctypes.create_string_buffer(8192)
        msb = (ctypes.c_float * 256)()
        cnt = (ctypes.c_int * 256)()
        n = lib.vc_profile_collect(names, 8192, msb, cnt, 256)
        lib.vc_profile_enable(0)
        fam_names = names.value.decode().split(",") if n else []
        work = family_work(p, B, n_params=n_params)
        for i, fn in enumerate(fam_names):
            families[fn] = {"ms_per_step": msb[i] / psteps, "launches_per_step": cnt[i] / psteps}
            if fn in work:
                kind, amount = work[fn]
                rate = amount / (msb[i] / psteps / 1e3)
                families[fn].update(bound=kind, achieved=rate / (1e12 if kind == "tensor" else 1e9))
        layer_fams = [k for k in families if k.startswith("conv") and k[4:5].isdigit()]
        if layer_fams:  # the 13 forward convolutions are one kernel family; per-layer entries stay for diagnosis
            ms_c = sum(families[k]["ms_per_step"] for k in layer_fams)
            families["conv"] = {"ms_per_step": ms_c, "launches_per_step": sum(families[k]["launches_per_step"] for k in layer_fams),
                                "bound": "tensor", "achieved": work["conv"][1] / (ms_c / 1e3) / 1e12}
        for agg, pre in (("conv_wgrad", "wgrad"), ("conv_dgrad", "dgrad")):
            parts = [k for k in families if k.startswith(pre) and k[len(pre):len(pre) + 1].isdigit()]
            if parts:
                ms_c = sum(families[k]["ms_per_step"] for k in parts)
                families[agg] = {"ms_per_step": ms_c, "launches_per_step": sum(families[k]["launches_per_step"] for k in parts),
                                 "bound": "tensor", "achieved": work[agg][1] / (ms_c / 1e3) / 1e12}
                layer_fams += parts
        # the LSTM 4-gate GEMMs as one group (north_star: "achieved fraction of the LSTM-GEMM roofline"): recurrence
        # forward + backward, weight gradient and input gradient over the time spent in those kernels
        lstm_parts = [k for k in ("lstm_fwd_seq", "lstm_fwd_step", "lstm_bwd_seq", "lstm_bwd_step", "lstm_wgrad", "lstm_dx")
                      if k in families and k in work]
        lstm_roofline = None
        if lstm_parts:
            ms_l = sum(families[k]["ms_per_step"] for k in lstm_parts)
            fl = sum(work[k][1] for k in lstm_parts)
            pk = peaks.get("bf16_tflops_sustained", 1590.0)
            lstm_roofline = {"kernels": lstm_parts, "bound": "tensor", "achieved": fl / (ms_l / 1e3) / 1e12, "peak": pk,
                             "unit": "TFLOP/s", "frac": fl / (ms_l / 1e3) / 1e12 / pk, "ms_per_step": ms_l}
        if families:
            top = max((k for k in families if k not in layer_fams), key=lambda k: families[k]["ms_per_step"])
            f = families[top]
            if "bound" in f:
                tensor = f["bound"] == "tensor"
                peak = peaks.get("bf16_tflops_sustained" if tensor else "hbm_gbs")
                src = "measured (MEASURED_PEAKS.json, %s)" % ("sustained bf16" if tensor else "copy")
                if peak is None:
                    peak, src = (1590.0 if tensor else 6650.0), "fallback"
                traffic = None
                try:  # profiles/traffic.json holds ncu DRAM bytes per launch of the DEFAULT workload only
                    if name == DEFAULT_WORKLOAD:
                        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(top)
                except Exception:
                    pass
                roofline = {"kernel": top, "bound": f["bound"], "achieved": f["achieved"], "peak": peak,
                            "unit": "TFLOP/s" if tensor else "GB/s", "frac": f["achieved"] / peak, "traffic": traffic,
                            "peak_source": src, "ms_per_launch": f["ms_per_step"] / max(f["launches_per_step"], 1),
                            "share_of_step": f["ms_per_step"] / sum(x["ms_per_step"] for k, x in families.items()
                                                                      if k not in layer_fams)}

    cpu_baseline = None
    if full and rank == 0 and world == 1 and not args.no_cpu_baseline:
        ra = argparse.Namespace(**vars(args))
        ra.steps, ra.warmup = 3, 1
        r = run_reference(ra, w, name)
        cpu_baseline = r["cpu_baseline"]

    line = None
    if rank == 0:
        peaks_tf = None
        try:
            peaks_tf = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops_sustained")
        except Exception:
            pass
        tf = total_flops(p, B, w["vgg"]) * world / (ms / steps / 1e3) / 1e12
        line = {"metric": "captions/sec (ELBO train step)", "value": value, "unit": "captions/s", "n_gpus": world,
                "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms / steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": name, "images_per_step_per_gpu": B, "captions_per_step_per_gpu": N,
                           "seq_len": T, "vocab": V, "prior": w["prior"], "on_device_vgg16_forward": w["vgg"],
                           "fine_tune": bool(w.get("fine_tune")), "c_v": w["c_v"],
                           "parallelism": "dp%d" % world,
                           "pipeline": ("frozen VGG16 forward of batch i+1 on a second stream overlaps the caption-model "
                                        "step of batch i (one forward and one step per timed step)") if pipelined else "none",
                           "l2_policy": "per-step working set (>= 0.3 GB logits + 80 MB weights/optimizer state) exceeds the 126 MB L2"},
                "e2e": {"value": e2e_val, "unit": "captions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 64,
                        "ms_per_step": ms_e2e / steps, "last_step": last,
                        "result_read": "every step's 64-byte result is read on the host inside the timed region, one step "
                                       "behind the launch (vc_step_result_queue / vc_step_result)"},
                "gpu_launches": launches, "clocks": clocks, "step_tflops": tf,
                "step_tensor_frac": (tf / world / peaks_tf) if peaks_tf else None}
        if world > 1:
            line.update(allreduce=allreduce, dp_check=dp_check, rec_loss_mean_over_ranks=loss_mean)
        if full:
            line.update(roofline=roofline, lstm_roofline=lstm_roofline, cpu_baseline=cpu_baseline,
                        serial_ms_per_step=(ms_serial / steps) if ms_serial else None, families=families)
    eng.close()
    del dev, host
    torch.cuda.empty_cache()
    return line


# BASELINE.json configs[2] / configs[3] ride along with the headline workload at every N: their
# per-GPU shapes, fewer steps, reported under the line's `configs` key
EXTRA_CONFIGS = ("cfg3_feats_gmm_cv_b128", "cfg4_finetune_ag_cv_b256")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--ref-batch", type=int, default=32, help="images per step of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (used under ncu only)")
    ap.add_argument("--no-extra-configs", action="store_true",
                    help="default workload only: do not also time BASELINE configs 3 and 4 (`configs` key)")
    args = ap.parse_args()
    quiet_stdout()
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            fn = run_decode_reference if w.get("decode") else run_reference
            emit(fn(args, w, args.workload))
        return 0
    if w.get("decode"):
        import torch
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback")
        return run_decode(args, w, args.workload, rank, world, local_rank)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    line = run_train(args, args.workload, rank, world, local_rank, args.steps, args.warmup, full=True)
    if args.workload == DEFAULT_WORKLOAD and not args.no_extra_configs and not args.no_e2e:
        extra = {}
        for name in EXTRA_CONFIGS:
            try:
                sub = run_train(args, name, rank, world, local_rank, max(3, min(args.steps, 10)), 3, full=False)
            except Exception as e:  # the headline line must survive a failing side workload
                sub = {"error": "%s: %s" % (type(e).__name__, e)}
                if world > 1:
                    raise
            if rank == 0:
                keep = ("value", "unit", "ms_per_step", "steps", "e2e", "allreduce", "dp_check", "rec_loss_mean_over_ranks",
                        "step_tflops", "step_tensor_frac", "gpu_launches", "error")
                extra[name] = {k: sub[k] for k in keep if k in sub}
                if "config" in sub:
                    extra[name]["config"] = {k: sub["config"][k] for k in ("images_per_step_per_gpu", "captions_per_step_per_gpu",
                                                                          "prior", "fine_tune", "c_v", "parallelism")}
        if rank == 0:
            line["configs"] = extra
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
