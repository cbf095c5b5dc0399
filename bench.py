#!/usr/bin/env python
"""bench.py — decode throughput of the TGIS continuous-batching hot path on B200 (contract in the task prompt / DESIGN.md).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference] [--no-extra]

Workloads (BASELINE.json):
  llama3-8b-gptq   the north-star target "batched int4 Llama-8B decode at bs=64": Llama-3-8B (GQA-8) GPTQ int4 g128, bs=64,
                   seq 1024->2048, tensor parallel over N GPUs.  DEFAULT AT EVERY N: it is the one int4 model that shards at
                   tp = 1/2/4/8, so the driver's 1/2/4/8 curve compares like with like.
  llama2-7b-gptq   config[2]: Llama-2-7B GPTQ int4, 1 GPU, bs=64, 1024->2048 (row-sharding stops at tp=2: 11008/tp % 128).
                   Measured in the same run at N=1 and reported under "extra".
  llama2-7b-fp16   config[3]: Llama-2-7B fp16, tensor parallel, bs=64.  Measured in the same run at N>1, under "extra".
  tinyllama-fp16   config[1]: TinyLlama-1.1B fp16, bs=32, 512->1024 (on request).
  llama3-70b-gptq-mixed   config[4]: Llama-3-70B int4, NUM_SHARD=8, continuous batching (Poisson arrivals, prompts U[256,1024], outputs
                   U[128,512], up to 128 running) through Prefill / NextToken with prune + concatenate; llama3-8b-gptq-mixed is the
                   same session on one GPU (on request).

A "step" is one decode step of the whole batch (one token per sequence).  The run really generates the whole trajectory
L0 -> L1 from the prompt, so KV contents, block tables and lengths are what the serving path produces:
  value : B*K / device time of K consecutive graph-replayed steps centred on the mean context (L0+L1)/2, inputs resident in
          HBM, no host round trip (CUDA events, max over ranks)
  e2e   : the same metric over ALL the other decode steps of the trajectory through the public API
          `FlashCausalLM.generate_token(batch)` driven from host buffers: every step copies the step's input token ids from
          pinned host memory and reads the chosen ids back to the host
  roofline : attn_decode_paged kernel, algorithmic KV bytes per launch / CUDA-event time per launch vs the measured HBM peak;
          "gemm" beside it: the int4 / fp16 linears' weight bytes per step / their CUDA-event time per step
  self_check : greedy ids of a 2-layer model of the same widths through the same code path vs the CPU oracle (tests' rule:
          exact outside the 2-ulp tie band)
  cpu_baseline : the reference's CPU CausalLM path (HF eager fp32, padded batch, greedy) on a bounded sample
`--impl reference` times that CPU path alone (rank 0 only) and prints the same JSON line with "impl": "reference".
Data is synthetic: seeded random weights of the named architecture, prompts "test " * L0 (the reference's own
load recipe, utils/memory_characterizer.py:219-240).
"""
from __future__ import annotations

import argparse
import ctypes
import gc
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (arch, quantize, batch, prompt_len, end_len)
    "llama3-8b-gptq": ("llama-3-8b", "gptq", 64, 1024, 2048),
    "llama2-7b-gptq": ("llama-2-7b", "gptq", 64, 1024, 2048),
    "llama2-7b-fp16": ("llama-2-7b", None, 64, 1024, 2048),
    "llama3-70b-gptq": ("llama-3-70b", "gptq", 128, 512, 1024),  # steady-state sibling of config[4]; needs --gpus 2 or more
    "pythia-12b-fp16": ("pythia-12b", None, 64, 1024, 2048),  # GPT-NeoX, the second flash family (parallel residual, partial rotary)
    "tiny-neox": ("tiny-neox", None, 4, 32, 256),
    "tinyllama-fp16": ("tinyllama-1.1b", None, 32, 512, 1024),
    "tiny-test": ("tiny-test", None, 4, 32, 256),
}
# continuous-batching sessions (BASELINE config[4] and a single-GPU sibling): (arch, quantize, max batch, prompt range, output range,
# requests, Poisson arrivals per second)
MIXED_WORKLOADS = {
    "llama3-70b-gptq-mixed": ("llama-3-70b", "gptq", 128, (256, 1024), (128, 512), 384, 60.0),
    "llama3-8b-gptq-mixed": ("llama-3-8b", "gptq", 128, (256, 1024), (128, 512), 384, 120.0),
    "tiny-test-mixed": ("tiny-test", None, 8, (8, 60), (4, 24), 40, 2000.0),
}
DEFAULT_WORKLOAD = "llama3-8b-gptq"
METRIC = "decode_tokens_per_s"
UNIT = "tokens/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe).  Started well before
    the timed region (nvidia-smi needs a second to enumerate an 8-GPU box); only samples between mark() and stop() count."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int = 0):
        self.lines = []
        self.proc = None
        self.index = index
        self.first = 0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self, wait_s: float = 5.0):
        """call right before the timed region: waits until nvidia-smi delivers, then forgets the earlier samples"""
        if self.proc is None:
            return
        t_end = time.time() + wait_s
        while not self.lines and time.time() < t_end:
            time.sleep(0.02)
        self.first = max(0, len(self.lines) - 1)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines[self.first:]:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ======================================================================================================
# CPU baseline: the reference's CausalLM path (models/causal_lm.py:548-739 on the hf_transformers engine) restated
# ======================================================================================================
def cpu_reference_decode(arch: str, B: int, ctx: int, steps: int, warmup: int, sample_layers: int):
    """Times greedy decode steps of the reference's CPU path on a bounded sample: `sample_layers` of the model's layers
    (same widths, fp32, HF eager attention, padded rectangular KV at context `ctx`), scaled to the full depth.
    -> dict(value tokens/s of the full model, cores, sample, ms_per_step full depth, steps timed per sample)."""
    from oracle import causal_lm as ocl  # oracle/ is the checker + CPU baseline only (never on the product path)
    return ocl.time_decode(arch, B, ctx, steps, warmup, sample_layers)


def run_reference(args, workload):
    arch, quantize, B, L0, L1 = WORKLOADS[workload]
    rank = int(os.getenv("RANK", "0"))
    if rank != 0:
        return 0
    ctx = (L0 + L1) // 2
    r = cpu_reference_decode(arch, B, ctx, args.steps, args.warmup, args.cpu_layers)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "arch": arch, "batch": B, "prompt_len": L0, "end_len": L1, "context": ctx,
                   "steps_timed_per_sample": r["steps_timed"], "extrapolated_in_depth": r["extrapolated"]},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if not args.no_extra and workload != "tiny-test":
        try:  # BASELINE config[0]: gpt2 CausalLM fp32 CPU, bs=4, 128 -> 256 (whole model, really run; random-init: no weights here)
            from oracle import causal_lm as ocl
            line["extra"] = {"config0_gpt2_cpu": ocl.time_decode_gpt2(4, 192, steps=8, warmup=2)}
        except Exception as e:  # noqa: BLE001
            line["extra"] = {"config0_gpt2_cpu": {"failed": f"{type(e).__name__}: {e}"}}
    print(json.dumps(line), flush=True)
    return 0


# ======================================================================================================
# GPU arm
# ======================================================================================================
def build_model(workload, world, rank, num_layers=None, cfg_override=None, weights=None, blocks=None):
    import torch
    import tgis_b200  # noqa: F401
    from tgis_b200.inference_engine import InferenceEngine
    from tgis_b200.models.flash_causal_lm import FlashCausalLM
    from tgis_b200.utils.dist import initialize_torch_distributed
    from tgis_b200.utils.synthetic import SyntheticWeights, make_tokenizer, model_config

    arch, quantize, B, L0, L1 = WORKLOADS[workload]
    cfg = cfg_override or model_config(arch, quantize=quantize, max_position_embeddings=max(4096, L1 + 64), num_layers=num_layers)
    local = int(os.getenv("LOCAL_RANK", rank))
    torch.cuda.set_device(local % torch.cuda.device_count())
    device = torch.device("cuda", torch.cuda.current_device())
    pg = initialize_torch_distributed(world, rank)
    if weights is None:
        weights = SyntheticWeights(cfg, device, torch.float16, pg, quantize=quantize)
    tok = make_tokenizer(cfg.vocab_size)
    engine = InferenceEngine("<synthetic>", None, torch.float16, quantize, cfg, L1, weights=weights, tokenizer=tok)
    if blocks is None:
        blocks = B * ((L1 + 16) // 16 + 1) + 8
    model = FlashCausalLM("<synthetic>", None, "tgis_native", torch.float16, quantize, cfg, engine=engine, num_kv_blocks=blocks)
    return model, cfg


def make_batch_pb(B, L0, n_new, batch_id=0):
    from tgis_b200 import pb
    reqs = []
    for i in range(B):
        reqs.append(pb.Request(
            id=i, inputs="test " * 10000, input_length=L0, truncate=True, max_output_length=n_new,
            parameters=pb.NextTokenChooserParameters(temperature=0.0, top_k=0, top_p=1.0, typical_p=0.0, min_new_tokens=n_new)))
    return pb.Batch(id=batch_id, requests=reqs, total_tokens=B * (L0 + n_new))


def algorithmic_bytes_per_step(cfg, quantize, B, ctx, tp):
    """SURVEY.md §8d formula, per GPU -> (whole step, attention KV read of one layer, linear weights incl. head)."""
    H, I, V, nl = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size, cfg.num_hidden_layers
    d = H // cfg.num_attention_heads
    h, kv = cfg.num_attention_heads, cfg.num_key_value_heads
    mlp_mats = 2 if getattr(cfg, "model_type", "llama") == "gpt_neox" else 3  # h_to_4h, 4h_to_h | gate, up, down
    lin_params = nl * ((h + 2 * kv) * d * H + h * d * H + mlp_mats * I * H)
    if quantize == "gptq":
        w_lin = lin_params * (0.5 + (2 + 0.5) / 128)
    else:
        w_lin = lin_params * 2
    w_head = V * H * 2
    kv_read = B * ctx * nl * 2 * kv * d * 2
    kv_write = B * nl * 2 * kv * d * 2
    logits = B * V * 2
    return (w_lin + w_head + kv_read + kv_write + logits) / tp, kv_read / tp / nl, (w_lin + w_head) / tp


def measure_workload(args, workload, world, rank, full: bool):
    """One workload, the whole L0 -> L1 trajectory.  `full`: also the per-kernel rooflines (eager passes) and prefill detail."""
    import torch
    import torch.distributed as dist

    arch, quantize, B, L0, L1 = WORKLOADS[workload]
    K, W = args.steps, args.warmup
    model, cfg = build_model(workload, world, rank, args.layers)
    from tgis_b200 import _lib
    lib = _lib.load()
    dev = model.device
    n_new = L1 - L0
    mid = (L0 + L1) // 2

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

    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()
    # ---------------------------------------------------------------- prefill (timed once, reported beside the metric)
    batch, errs = model.batch_type.from_pb(make_batch_pb(B, L0, n_new), model.tokenizer, model.dtype, dev, None, None, True)
    assert not errs
    e2e_ms, e2e_steps, e2e_ctx_sum = 0.0, 0, 0.0
    host_ids = torch.empty(B, dtype=torch.int64).pin_memory()
    with torch.inference_mode():
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        toks = model.generate_token(batch, first=True)[0]
        e1.record()
        barrier()
        prefill_ms = max_over_ranks(e0.elapsed_time(e1))
        cur = L0  # tokens cached before the next decode step (that step attends over cur + 1)

        def e2e_segment(n_steps, toks, cur):
            """n_steps decode steps through generate_token with host buffers in the timed region; -> (ms, toks, cur)"""
            if n_steps <= 0:
                return 0.0, toks, cur
            barrier()
            t0 = time.perf_counter()
            e0.record()
            for _ in range(n_steps):
                host_ids.copy_(torch.tensor([t.token_id for t in toks], dtype=torch.int64))
                batch.input_ids.copy_(host_ids, non_blocking=True)      # H2D: this step's input token ids
                toks = model.generate_token(batch)[0]                    # D2H: the chosen ids (one read per step)
            e1.record()
            barrier()
            wall = (time.perf_counter() - t0) * 1e3
            return max_over_ranks(max(e0.elapsed_time(e1), wall)), toks, cur + n_steps

        # ---- e2e, first half of the trajectory (3 untimed steps first: the fused state and its CUDA graph come to exist)
        for _ in range(3):
            toks = model.generate_token(batch)[0]
            cur += 1
        n_eager = K if full else 0
        start_a = mid - (K + n_eager) // 2 - W
        seg_steps = max(0, start_a - cur)
        ctx0 = cur
        ms, toks, cur = e2e_segment(seg_steps, toks, cur)
        e2e_ms += ms
        e2e_steps += seg_steps
        e2e_ctx_sum += seg_steps * (ctx0 + (seg_steps + 1) / 2.0)

        # ---- device-resident window around the mean context: no host round trip inside the timed region
        st = batch._fused
        batch.input_ids.copy_(torch.tensor([t.token_id for t in toks], dtype=torch.int64))

        def device_step(use_graph=True):
            model._run_fused_step(batch, st, use_graph)  # chains the chosen ids into input_ids on the device

        for _ in range(W):
            device_step()
            cur += 1
        attn_ms = gemm_ms = None
        n_attn = n_gemm = 0
        ctx_roof = None
        if full:
            # eager steps carry the per-kernel events (events are not captured into graphs); two kernel families, K/2 steps each
            which_gemm = 2 if quantize == "gptq" else 3  # B200_TIME_GEMM_W4A16 / B200_TIME_GEMM_F16
            halves = [(1, K - K // 2), (which_gemm, K // 2)]
            for which, n_steps in halves:
                timing = lib.b200_timing_create(n_steps * (4 * cfg.num_hidden_layers + 2) + 8)
                lib.b200_timing_attach(timing, which)
                barrier()
                for _ in range(n_steps):
                    device_step(use_graph=False)
                barrier()
                lib.b200_timing_attach(None, 0)
                tot = ctypes.c_float(0)
                n_timed = lib.b200_timing_collect(timing, ctypes.byref(tot))
                if n_timed < 0:
                    raise RuntimeError(lib.b200_last_error().decode())
                lib.b200_timing_destroy(timing)
                if which == 1:
                    attn_ms, n_attn = tot.value / max(n_timed, 1), n_timed
                    ctx_roof = cur + (n_steps + 1) / 2.0
                else:
                    gemm_ms, n_gemm = tot.value / max(n_steps, 1), n_timed  # per step
                cur += n_steps
        # headline pass
        launches0 = lib.b200_launch_count()
        sampler.mark()
        barrier()
        e0.record()
        for _ in range(K):
            device_step()
        e1.record()
        barrier()
        dev_ms = max_over_ranks(e0.elapsed_time(e1))
        clocks = sampler.stop()
        ctx_value = cur + (K + 1) / 2.0
        cur += K
        # the host mirrors of the lengths advance too (generate_token was bypassed for these steps)
        skipped = W + n_eager + K
        for i in range(B):
            batch.input_lengths[i] += skipped
            batch.next_token_chooser.current_tokens[i] += skipped
        batch.max_seqlen += skipped
        # graph replays launch the same kernels as an eager step; count them from one eager step's counter delta
        l0 = lib.b200_launch_count()
        device_step(use_graph=False)
        torch.cuda.synchronize()
        launches_per_step = lib.b200_launch_count() - l0
        cur += 1
        for i in range(B):
            batch.input_lengths[i] += 1
            batch.next_token_chooser.current_tokens[i] += 1
        batch.max_seqlen += 1

        # ---- e2e, second half of the trajectory: up to the last token the requests asked for
        toks = model.generate_token(batch)[0]
        cur += 1
        seg_steps = max(0, (L1 - 1) - cur)
        ctx0 = cur
        ms, toks, cur = e2e_segment(seg_steps, toks, cur)
        e2e_ms += ms
        e2e_steps += seg_steps
        e2e_ctx_sum += seg_steps * (ctx0 + (seg_steps + 1) / 2.0)

    hbm_peak, peak_kind = peaks()
    step_bytes, _, w_bytes = algorithmic_bytes_per_step(cfg, quantize, B, ctx_value, world)
    value = B * K / (dev_ms / 1e3)
    res = {
        "workload": workload, "value": value, "ms_per_step": dev_ms / K, "mean_context_timed": ctx_value,
        "step_roofline": {"algorithmic_bytes_per_step_per_gpu": step_bytes, "hbm_gbs_achieved": step_bytes / (dev_ms / K / 1e3) / 1e9,
                          "frac_of_hbm_peak": step_bytes / (dev_ms / K / 1e3) / 1e9 / hbm_peak, "peak_kind": peak_kind},
        "prefill": {"tokens": B * L0, "ms": prefill_ms, "tokens_per_s": B * L0 / (prefill_ms / 1e3)},
        "e2e": {"value": B * e2e_steps / (e2e_ms / 1e3) if e2e_ms > 0 else None, "unit": UNIT, "h2d_bytes_per_step": B * 8,
                "d2h_bytes_per_step": B * 8, "ms_per_step": e2e_ms / max(e2e_steps, 1), "steps": e2e_steps,
                "mean_context": e2e_ctx_sum / max(e2e_steps, 1), "api": "FlashCausalLM.generate_token",
                "trajectory": f"every decode step {L0}->{L1} outside the {skipped + 5}-step device-timed window"},
        "gpu_launches": int(launches_per_step * K), "launches_per_step": int(launches_per_step), "clocks": clocks,
        "arch": arch, "quantize": quantize, "batch": B, "prompt_len": L0, "end_len": L1, "layers": cfg.num_hidden_layers,
    }
    if full and attn_ms:
        _, attn_bytes_roof, _ = algorithmic_bytes_per_step(cfg, quantize, B, ctx_roof, world)
        achieved = attn_bytes_roof / (attn_ms / 1e3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            try:
                with open(tpath) as f:
                    traffic = json.load(f).get(workload, {}).get("attn_decode_traffic_bytes_per_launch")
            except Exception:
                traffic = None
        res["roofline"] = {"kernel": "attn_decode_paged_kernel", "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                           "frac": achieved / hbm_peak, "traffic": traffic, "peak_kind": peak_kind,
                           "algorithmic_bytes_per_launch": attn_bytes_roof, "avg_launch_ms": attn_ms, "launches_timed": n_attn,
                           "share_of_step": attn_ms * cfg.num_hidden_layers / (dev_ms / K),
                           "timing": "CUDA events around each eager launch (adds the event gaps: graph-replayed launches are shorter, "
                                     "see profiles/ step traces)"}
        if gemm_ms:
            res["roofline"]["gemm"] = {"kernel": "gemm_w4a16_kernel" if quantize == "gptq" else "gemm_f16_kernel", "bound": "hbm",
                                       "algorithmic_weight_bytes_per_step": w_bytes, "ms_per_step": gemm_ms,
                                       "achieved": w_bytes / (gemm_ms / 1e3) / 1e9, "unit": "GB/s",
                                       "frac": w_bytes / (gemm_ms / 1e3) / 1e9 / hbm_peak, "share_of_step": gemm_ms / (dev_ms / K),
                                       "launches_timed": n_gemm}
    # free everything before the next workload
    del batch, st, model
    gc.collect()
    torch.cuda.empty_cache()
    return res


def run_mixed(args, workload):
    """BASELINE config[4]: continuous batching with add-on prefills, driven through the shard's own Prefill / NextToken messages
    by tools/router_sim.py (the Rust router's loop restated).  Decode-step and prefill throughput are reported separately."""
    import torch
    import torch.distributed as dist

    from tools import router_sim

    world = int(os.getenv("WORLD_SIZE", "1"))
    rank = int(os.getenv("RANK", "0"))
    arch, quantize, max_bs, prompt_range, new_range, n_req, rate = MIXED_WORKLOADS[workload]
    n_req = args.requests or n_req
    max_len = prompt_range[1] + new_range[1]
    WORKLOADS[workload] = (arch, quantize, max_bs, prompt_range[1], max_len)
    blocks = max_bs * ((max_len + 16) // 16 + 1) + 8
    model, cfg = build_model(workload, world, rank, args.layers, blocks=blocks)
    from tgis_b200 import _lib, pb
    from tgis_b200.server import Cache, TextGenerationService
    lib = _lib.load()
    service = TextGenerationService(model, Cache(), ["unix:///dev/null"])
    requests = router_sim.make_requests(n_req, rate, prompt_range, new_range, cfg.vocab_size, seed=0, filler=3)  # "test " * L
    budget = max_bs * max_len
    bytes_acc = {"decode": 0.0}
    H, I, V, nl = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size, cfg.num_hidden_layers
    d = H // cfg.num_attention_heads

    def on_step(bs, ctx_sum):
        step_bytes, _, _ = algorithmic_bytes_per_step(cfg, quantize, 1, 0, world)        # weights + head (+ 1 row of logits)
        kv = (ctx_sum + bs) * nl * 2 * cfg.num_key_value_heads * d * 2 / world              # KV read of every context (+ the new token)
        bytes_acc["decode"] += step_bytes + kv + (bs - 1) * V * 2 / world

    def sync():
        torch.cuda.synchronize()

    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()
    l0 = lib.b200_launch_count()
    if world > 1:
        dist.barrier()
    sampler.mark()
    def agree(dt):
        """every rank advances the session clock by the slowest rank's time: identical admission decisions on all shards"""
        if world == 1:
            return dt
        t = torch.tensor([dt], dtype=torch.float64, device=model.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    with torch.inference_mode():
        out = router_sim.run_session(service, pb, requests, max_bs, lambda p: "test " * len(p), max_batch_tokens=budget, sync=sync,
                                     on_decode_step=on_step, agree=agree)
    clocks = sampler.stop()
    st = out["stats"]
    launches = lib.b200_launch_count() - l0
    hbm_peak, peak_kind = peaks()

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=model.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    decode_s, prefill_s = maxr(st["decode_s"]), maxr(st["prefill_s"])
    value = st["decode_tokens"] / decode_s
    assert all(len(out["tokens"][r.id]) == r.max_new for r in requests), "a request did not receive all its tokens"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": st["decode_steps"], "warmup": 0,
        "ms_per_step": decode_s / max(st["decode_steps"], 1) * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f16" if quantize is None else "int4 weights x f16 activations, f32 accumulate", "data": "synthetic",
        "config": {"workload": workload, "arch": arch, "quantize": quantize, "max_batch": max_bs, "prompt_len": list(prompt_range),
                   "new_tokens": list(new_range), "requests": n_req, "poisson_arrivals_per_s": rate, "seed": 0, "kv_block": 16,
                   "parallelism": f"tp{world}", "layers": cfg.num_hidden_layers,
                   "driver": "tools/router_sim.py: add-on Prefill (to_prune) + NextToken over every cached batch id, as router/src/batcher.rs",
                   "timing": "wall clock around each RPC, device synchronised on both sides (host bookkeeping of prune / concatenate included)"},
        "tokens_per_s_per_gpu": value / world,
        "decode": {"tokens": st["decode_tokens"], "seconds": decode_s, "steps": st["decode_steps"],
                   "mean_batch": st["batch_size_sum"] / max(st["decode_steps"], 1), "max_batch": st["max_batch"],
                   "steps_after_a_concatenate": st["concat_steps"]},
        "prefill": {"tokens": st["prefill_tokens"], "ms": prefill_s * 1e3, "calls": st["prefill_calls"],
                    "tokens_per_s": st["prefill_tokens"] / prefill_s if prefill_s > 0 else None},
        "step_roofline": {"algorithmic_bytes_per_gpu_all_decode_steps": bytes_acc["decode"],
                          "hbm_gbs_achieved": bytes_acc["decode"] / decode_s / 1e9,
                          "frac_of_hbm_peak": bytes_acc["decode"] / decode_s / 1e9 / hbm_peak, "peak_kind": peak_kind},
        "e2e": {"value": (st["decode_tokens"] + st["prefill_calls"]) / (decode_s + prefill_s), "unit": UNIT,
                "note": "generated tokens / (prefill + decode time): the session as the router sees it, host buffers in and out of every RPC",
                "h2d_bytes_per_step": None, "d2h_bytes_per_step": None},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.cuda.synchronize()
        sys.stdout.flush()
        try:
            store = dist.distributed_c10d._get_default_store()
            store.add("bench_done", 1)
            t_end = time.time() + 30
            while int(store.add("bench_done", 0)) < world and time.time() < t_end:
                time.sleep(0.05)
        except Exception:
            time.sleep(1.0)
        os._exit(0)
    return 0


def self_check(arch: str, quantize):
    """2 layers of the workload's widths (small vocabulary), ragged prompts, greedy decode through from_pb / generate_token
    (prefill, eager steps, the CUDA-graph-replayed fused step) vs the CPU oracle; exact outside the 2-ulp tie band."""
    import tempfile

    import torch
    from safetensors.torch import save_file

    from oracle import llama as oll  # checker only
    from tgis_b200 import pb
    from tgis_b200.inference_engine import InferenceEngine
    from tgis_b200.models.flash_causal_lm import FlashCausalLM
    from tgis_b200.utils.dist import FakeGroup
    from tgis_b200.utils.synthetic import ARCHS, llama_config, make_tokenizer
    from tgis_b200.utils.weights import Weights

    H, I, _, h, kv, _ = ARCHS[arch]
    cfg = llama_config(arch, quantize=quantize, max_position_embeddings=512, num_layers=2)
    cfg.vocab_size = 4096
    ocfg = oll.LlamaConfig(H, I, 2, h, kv, cfg.vocab_size, cfg.rms_norm_eps, cfg.rope_theta)
    sd = oll.make_state_dict(ocfg, seed=11, quantize=quantize, std=0.02)
    dev = "cuda:0"
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "model.safetensors")
        save_file({k: v.contiguous() for k, v in sd.items()}, path)
        weights = Weights([path], device=dev, dtype=torch.float16, process_group=FakeGroup(0, 1))
        tok = make_tokenizer(cfg.vocab_size)
        engine = InferenceEngine(tmp, None, torch.float16, quantize, cfg, 512, weights=weights, tokenizer=tok)
        model = FlashCausalLM(tmp, None, "tgis_native", torch.float16, quantize, cfg, engine=engine, num_kv_blocks=128)
    g = torch.Generator().manual_seed(0)
    lens = (40, 64, 17, 33, 48, 5, 64, 21)
    prompts = [torch.randint(4, cfg.vocab_size, (L,), generator=g).tolist() for L in lens]
    n_new = 8
    reqs = [pb.Request(id=i, inputs=" ".join(f"<tok{t}>" for t in p), input_length=len(p), max_output_length=n_new,
                       parameters=pb.NextTokenChooserParameters(temperature=0.0, top_p=1.0, min_new_tokens=n_new))
            for i, p in enumerate(prompts)]
    got = [[] for _ in prompts]
    with torch.inference_mode():
        batch, errs = model.batch_type.from_pb(pb.Batch(id=0, requests=reqs), tok, torch.float16, model.device, None, None, True)
        assert not errs
        out = model.generate_token(batch, first=True)
        for _ in range(n_new - 1):
            for t in out[0]:
                got[t.request_id].append(t.token_id)
            out = model.generate_token(batch)
        for t in out[0]:
            got[t.request_id].append(t.token_id)
    torch.cuda.synchronize()
    oracle = oll.LlamaOracle(oll.build_shards(ocfg, sd, 1))
    ref, ref_logits = oracle.generate_greedy(prompts, n_new, banned_token=cfg.eos_token_id)
    checked = mismatches = ties = 0
    for b in range(len(prompts)):
        for step in range(n_new):
            lg = ref_logits[step][b].float().clone()
            lg[cfg.eos_token_id] = float("-inf")
            top2 = lg.topk(2).values
            if float(top2[0] - top2[1]) <= 2 * max(abs(float(top2[0])), 1.0) * 2.0 ** -10:
                ties += 1
                break  # after a tie the continuations may legitimately differ
            checked += 1
            if got[b][step] != int(ref[b, step]):
                mismatches += 1
                break
    del model, batch
    gc.collect()
    torch.cuda.empty_cache()
    return {"model": f"{arch} widths, 2 layers, vocab 4096, {quantize or 'fp16'}", "sequences": len(prompts), "new_tokens": n_new,
            "tokens_checked": checked, "ties_skipped": ties, "mismatches_outside_tie_band": mismatches,
            "steps_through_cuda_graph": max(0, n_new - 1 - 2)}


def run_gpu(args, workload):
    import torch
    import torch.distributed as dist

    world = int(os.getenv("WORLD_SIZE", "1"))
    rank = int(os.getenv("RANK", "0"))
    main = measure_workload(args, workload, world, rank, full=True)
    arch, quantize, B, L0, L1 = WORKLOADS[workload]
    line = {
        "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f16" if quantize is None else "int4 weights x f16 activations, f32 accumulate", "data": "synthetic",
        "config": {"workload": workload, "arch": arch, "quantize": quantize, "batch": B, "prompt_len": L0, "end_len": L1,
                   "mean_context_timed": main["mean_context_timed"], "kv_block": 16, "parallelism": f"tp{world}",
                   "l2_policy": "inputs larger than L2 (KV + weights per step >> 126 MB)", "layers": main["layers"]},
        "tokens_per_s_per_gpu": main["value"] / world,
        "step_roofline": main["step_roofline"], "prefill": main["prefill"], "e2e": main["e2e"],
        "gpu_launches": main["gpu_launches"], "clocks": main["clocks"],
    }
    if "roofline" in main:
        line["roofline"] = main["roofline"]
    extra = {}
    if not args.no_extra and args.workload is None:
        other = "llama2-7b-gptq" if world == 1 else "llama2-7b-fp16"  # BASELINE config[2] / config[3]
        try:
            r = measure_workload(args, other, world, rank, full=False)
            extra[other] = {k: r[k] for k in ("value", "ms_per_step", "mean_context_timed", "step_roofline", "prefill", "e2e", "clocks",
                                              "launches_per_step", "arch", "quantize", "batch", "prompt_len", "end_len")}
            extra[other]["parallelism"] = f"tp{world}"
        except Exception as e:  # noqa: BLE001
            extra[other] = {"failed": f"{type(e).__name__}: {e}"}
    if rank == 0:
        from tgis_b200.utils.synthetic import ARCHS as LLAMA_ARCHS
        llama_family = arch in LLAMA_ARCHS  # the self-check and the CPU arm restate the Llama graph
        if world == 1 and not args.no_extra and llama_family:
            try:
                line["self_check"] = self_check(arch, quantize)
            except Exception as e:  # noqa: BLE001
                line["self_check"] = {"failed": f"{type(e).__name__}: {e}"}
        if not args.no_cpu_baseline and world == 1 and llama_family:
            try:
                r = cpu_reference_decode(arch, B, (L0 + L1) // 2, max(2, min(4, args.steps)), 1, args.cpu_layers)
                line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                        "sample": f"failed: {type(e).__name__}: {e}"}
        if extra:
            line["extra"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        # collectives live inside the captured CUDA graphs: tearing the communicator down while they exist can hang, so
        # rendezvous on the host-side store instead of an NCCL barrier, then leave without destroy_process_group()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        try:
            store = dist.distributed_c10d._get_default_store()
            store.add("bench_done", 1)
            t_end = time.time() + 30
            while int(store.add("bench_done", 0)) < world and time.time() < t_end:
                time.sleep(0.05)
        except Exception:
            time.sleep(1.0)
        os._exit(0)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=48)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=list(WORKLOADS) + list(MIXED_WORKLOADS))
    ap.add_argument("--requests", type=int, default=None, help="mixed workloads: number of requests in the session")
    ap.add_argument("--layers", type=int, default=None, help="debug: override the number of layers")
    ap.add_argument("--cpu-layers", type=int, default=2, help="layers in the CPU baseline's bounded sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the second workload, the self-check and the gpt2 CPU line")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    workload = args.workload or DEFAULT_WORKLOAD
    if workload in MIXED_WORKLOADS:
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "the continuous-batching session is a GPU-arm workload; the CPU arm times "
                                                                  "single decode steps (use a decode workload)"}))
            return 0
        return run_mixed(args, workload)
    if args.impl == "reference":
        return run_reference(args, workload)
    return run_gpu(args, workload)


if __name__ == "__main__":
    sys.exit(main())
