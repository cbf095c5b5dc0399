#!/usr/bin/env python
"""bench.py — decode throughput of the TGIS continuous-batching hot path on B200 (contract in the task prompt / DESIGN.md).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

Workloads (BASELINE.json configs):
  llama2-7b-gptq   config[2]: Llama-2-7B GPTQ int4 g128, 1 GPU, bs=64, seq 1024->2048   (default at N=1: the config the
                   metric "decode tokens/sec/GPU (bs=64, seq 1k->2k)" is quoted on)
  llama2-7b-fp16   config[3]: Llama-2-7B fp16, tensor parallel over N GPUs, bs=64, 1024->2048 (default at N>1; int4
                   Llama-2-7B cannot be row-sharded beyond tp=2: 11008/tp is not a multiple of the group size 128)
  tinyllama-fp16   config[1]: TinyLlama-1.1B fp16, bs=32, 512->1024
  llama3-8b-gptq   north-star "Llama-8B": Llama-3-8B (GQA-8) GPTQ int4, bs=64, 1024->2048

A "step" is one decode step of the whole batch (one token per sequence).  The timed window is K consecutive steps
centred on the mean context of the workload (L = (start+end)/2), reached by really running prefill + decode from the
prompt, so KV contents, block tables and lengths are what the serving path produces.
  value : B*K / device time of K steps with all inputs resident in HBM (CUDA events, max over ranks)
  e2e   : same metric through the public API `FlashCausalLM.generate_token(batch)` driven from host buffers: every step
          copies the step's input token ids from pinned host memory and reads the chosen ids back to the host
  roofline : attn_decode_paged kernel, algorithmic KV bytes per launch / CUDA-event time per launch vs measured HBM peak
  cpu_baseline : the reference's CPU CausalLM path (HF eager fp32, padded batch, greedy) on a bounded sample
`--impl reference` times that CPU path alone (rank 0 only) and prints the same JSON line with "impl": "reference".
Data is synthetic: seeded random weights of the named architecture, prompts "test " * L0 (the reference's own
load recipe, utils/memory_characterizer.py:219-240).
"""
from __future__ import annotations

import argparse
import ctypes
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
    "llama2-7b-gptq": ("llama-2-7b", "gptq", 64, 1024, 2048),
    "llama2-7b-fp16": ("llama-2-7b", None, 64, 1024, 2048),
    "tinyllama-fp16": ("tinyllama-1.1b", None, 32, 512, 1024),
    "llama3-8b-gptq": ("llama-3-8b", "gptq", 64, 1024, 2048),
    "tiny-test": ("tiny-test", None, 4, 32, 256),
}
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
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int = 0):
        self.lines = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
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
    Returns (tokens_per_s_full_model, cores, sample_description, ms_per_step_sample)."""
    import torch
    from oracle import causal_lm as ocl  # oracle/ is the checker + CPU baseline only (never on the product path)
    return ocl.time_decode(arch, B, ctx, steps, warmup, sample_layers)


def run_reference(args, workload):
    arch, quantize, B, L0, L1 = WORKLOADS[workload]
    rank = int(os.getenv("RANK", "0"))
    if rank != 0:
        return 0
    ctx = (L0 + L1) // 2
    tps, cores, sample, ms = cpu_reference_decode(arch, B, ctx, args.steps, args.warmup, args.cpu_layers)
    line = {
        "impl": "reference", "metric": METRIC, "value": tps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "arch": arch, "batch": B, "prompt_len": L0, "end_len": L1, "context": ctx},
        "cpu_baseline": {"value": tps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": tps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ======================================================================================================
# GPU arm
# ======================================================================================================
def build_model(workload, world, rank, num_layers=None):
    import torch
    import tgis_b200  # noqa: F401
    from tgis_b200.inference_engine import InferenceEngine
    from tgis_b200.models.flash_causal_lm import FlashCausalLM
    from tgis_b200.utils.dist import initialize_torch_distributed
    from tgis_b200.utils.synthetic import SyntheticWeights, llama_config, make_tokenizer

    arch, quantize, B, L0, L1 = WORKLOADS[workload]
    cfg = llama_config(arch, quantize=quantize, max_position_embeddings=max(4096, L1 + 64), num_layers=num_layers)
    local = int(os.getenv("LOCAL_RANK", rank))
    torch.cuda.set_device(local % torch.cuda.device_count())
    device = torch.device("cuda", torch.cuda.current_device())
    pg = initialize_torch_distributed(world, rank)
    weights = SyntheticWeights(cfg, device, torch.float16, pg, quantize=quantize)
    tok = make_tokenizer(cfg.vocab_size)
    engine = InferenceEngine("<synthetic>", None, torch.float16, quantize, cfg, L1, weights=weights, tokenizer=tok)
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
    """SURVEY.md §8d formula, per GPU."""
    H, I, V, nl = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size, cfg.num_hidden_layers
    d = H // cfg.num_attention_heads
    h, kv = cfg.num_attention_heads, cfg.num_key_value_heads
    lin_params = nl * ((h + 2 * kv) * d * H + h * d * H + 2 * I * H + I * H)
    if quantize == "gptq":
        w_lin = lin_params * (0.5 + (2 + 0.5) / 128)
    else:
        w_lin = lin_params * 2
    w_head = V * H * 2
    kv_read = B * ctx * nl * 2 * kv * d * 2
    kv_write = B * nl * 2 * kv * d * 2
    logits = B * V * 2
    return (w_lin + w_head + kv_read + kv_write + logits) / tp, kv_read / tp / nl


def run_gpu(args, workload):
    import torch
    import torch.distributed as dist

    world = int(os.getenv("WORLD_SIZE", "1"))
    rank = int(os.getenv("RANK", "0"))
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

    # ---------------------------------------------------------------- prefill (timed once, reported beside the metric)
    batch, errs = model.batch_type.from_pb(make_batch_pb(B, L0, n_new), model.tokenizer, model.dtype, dev, None, None, True)
    assert not errs
    with torch.inference_mode():
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        model.generate_token(batch, first=True)
        e1.record()
        barrier()
        prefill_ms = max_over_ranks(e0.elapsed_time(e1))

        # ------------------------------------------------------------ run the real trajectory up to the timed window
        # phase A (device-resident "value"): window of K steps ending at context mid; phase B (e2e): K steps from mid.
        start_a = mid - K - W
        cur = L0  # tokens cached before the next decode step (that step attends over cur + 1)
        while cur < start_a:
            model.generate_token(batch)
            cur += 1
        # ---- phase A: device-resident loop through the fused step (no host round trip inside the timed region)
        st_launch0 = None
        for _ in range(3):  # make sure the fused state (and CUDA graph) of this batch exists
            model.generate_token(batch)
            cur += 1
        st = batch._fused
        kv = batch.past_key_values

        def device_step(use_graph=True):
            model._run_fused_step(batch, st, use_graph)
            batch.position_ids += 1
            batch.input_ids.copy_(st["next_ids"])

        for _ in range(W):
            device_step()
            cur += 1
        timing = lib.b200_timing_create(K * cfg.num_hidden_layers + 8)
        # eager steps carry the per-kernel events; graph replays cannot (events are not captured), so the roofline
        # pass runs the same K steps eagerly first, then the headline pass replays the graph
        lib.b200_timing_attach(timing, 1)
        barrier()
        for _ in range(K):
            device_step(use_graph=False)
        barrier()
        lib.b200_timing_attach(None, 0)
        tot = ctypes.c_float(0)
        n_timed = lib.b200_timing_collect(timing, ctypes.byref(tot))
        if n_timed < 0:
            raise RuntimeError(lib.b200_last_error().decode())
        attn_ms = tot.value / max(n_timed, 1)
        ctx_roof = cur + (K + 1) / 2.0  # mean context (incl. the token written) over those K steps
        cur += K
        # headline pass
        sampler = ClockSampler(torch.cuda.current_device())
        sampler.start()
        launches0 = lib.b200_launch_count()
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
        launches_eager_equiv = None
        # the host mirrors of the lengths advance too (generate_token was bypassed for 2K+W steps)
        for i in range(B):
            batch.input_lengths[i] += 2 * K + W
        batch.max_seqlen += 2 * K + W
        batch.cu_seqlens.add_(batch.cu_seqlens_q * (2 * K + W))
        for i in range(B):
            batch.next_token_chooser.current_tokens[i] += 2 * K + W

        # ---- phase B: end to end through generate_token with host buffers
        host_ids = torch.empty(B, dtype=torch.int64).pin_memory()
        toks = model.generate_token(batch)[0]
        cur += 1
        for _ in range(W):
            toks = model.generate_token(batch)[0]
            cur += 1
        barrier()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(K):
            host_ids.copy_(torch.tensor([t.token_id for t in toks], dtype=torch.int64))
            batch.input_ids.copy_(host_ids, non_blocking=True)      # H2D: this step's input token ids
            toks = model.generate_token(batch)[0]                    # D2H: the chosen ids (one read per step)
        e1.record()
        barrier()
        e2e_wall_ms = (time.perf_counter() - t0) * 1e3
        e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), e2e_wall_ms))
        ctx_e2e = cur + (K + 1) / 2.0
        cur += K

    # graph replays launch the same kernels as an eager step; count them from one eager step's counter delta
    l0 = lib.b200_launch_count()
    with torch.inference_mode():
        device_step(use_graph=False)
        torch.cuda.synchronize()
    launches_per_step = lib.b200_launch_count() - l0

    hbm_peak, peak_kind = peaks()
    step_bytes, attn_bytes = algorithmic_bytes_per_step(cfg, quantize, B, ctx_value, world)
    _, attn_bytes_roof = algorithmic_bytes_per_step(cfg, quantize, B, ctx_roof, world)
    value = B * K / (dev_ms / 1e3)
    e2e_value = B * K / (e2e_ms / 1e3)
    achieved = attn_bytes_roof / (attn_ms / 1e3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                traffic = json.load(f).get(workload, {}).get("attn_decode_traffic_bytes_per_launch")
        except Exception:
            traffic = None

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f16" if quantize is None else "int4 weights x f16 activations, f32 accumulate", "data": "synthetic",
        "config": {"workload": workload, "arch": arch, "quantize": quantize, "batch": B, "prompt_len": L0, "end_len": L1,
                   "mean_context_timed": ctx_value, "kv_block": 16, "parallelism": f"tp{world}",
                   "l2_policy": "inputs larger than L2 (KV + weights per step >> 126 MB)",
                   "layers": cfg.num_hidden_layers},
        "tokens_per_s_per_gpu": value / world,
        "step_roofline": {"algorithmic_bytes_per_step_per_gpu": step_bytes, "hbm_gbs_achieved": step_bytes / (dev_ms / K / 1e3) / 1e9,
                          "frac_of_hbm_peak": step_bytes / (dev_ms / K / 1e3) / 1e9 / hbm_peak, "peak_kind": peak_kind},
        "prefill": {"tokens": B * L0, "ms": prefill_ms, "tokens_per_s": B * L0 / (prefill_ms / 1e3)},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * 8, "d2h_bytes_per_step": B * 8,
                "ms_per_step": e2e_ms / K, "mean_context": ctx_e2e, "api": "FlashCausalLM.generate_token"},
        "gpu_launches": int(launches_per_step * K),
        "clocks": clocks,
        "roofline": {"kernel": "attn_decode_paged_kernel", "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "traffic": traffic, "peak_kind": peak_kind,
                     "algorithmic_bytes_per_launch": attn_bytes_roof, "avg_launch_ms": attn_ms, "launches_timed": n_timed,
                     "share_of_step": attn_ms * cfg.num_hidden_layers / (dev_ms / K)},
    }
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            try:
                tps, cores, sample, ms = cpu_reference_decode(arch, B, mid, max(2, min(4, K)), 1, args.cpu_layers)
                line["cpu_baseline"] = {"value": tps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                        "sample": f"failed: {type(e).__name__}: {e}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        # NCCL kernels live inside the captured CUDA graphs: tearing the communicator down while they exist can hang, so
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
    ap.add_argument("--workload", default=None, choices=list(WORKLOADS))
    ap.add_argument("--layers", type=int, default=None, help="debug: override the number of layers")
    ap.add_argument("--cpu-layers", type=int, default=2, help="layers in the CPU baseline's bounded sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    world = int(os.getenv("WORLD_SIZE", "1"))
    workload = args.workload or ("llama2-7b-gptq" if max(args.gpus, world) == 1 else "llama2-7b-fp16")
    if args.impl == "reference":
        return run_reference(args, workload)
    return run_gpu(args, workload)


if __name__ == "__main__":
    sys.exit(main())
