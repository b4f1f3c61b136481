#!/usr/bin/env python
"""Headline benchmark: env steps/s of the batched TensorGame transition.

Workload (BASELINE.json configs[1]): 3x3 matmul tensor (9x9x9), 2^20 parallel
games per GPU, coefficients {-2..2}.  One "step" = one tg_step launch over the
whole batch (every game makes one transition).  See DESIGN.md "Measurement".

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  For N>1 launch with torch.distributed.run.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

ALGO_BYTES = {4: 145, 9: 1490, 16: 8245}  # 2*S^3 + 3S + 1 + 4 per env step
FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback


def parse() -> argparse.Namespace:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=9, choices=[4, 9, 16], help="dim_3d S")
    ap.add_argument("--games", type=int, default=1 << 20, help="games per GPU")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--ctas-per-sm", type=int, default=0, help="tuning sweep only")
    ap.add_argument("--variant", type=int, default=0, help="tuning sweep only")
    return ap.parse_args()


def hbm_peak() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            self.path = tempfile.NamedTemporaryFile(prefix="clocks_", suffix=".csv", delete=False).name
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.index)],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.proc.wait()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.path)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _time_ms(fn, iters: int, torch) -> float:
    for _ in range(3):  # first launches carry host-side module-load latency
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def measure_extras(env, dev, S, shift, slab, tape3, R, values, probs, world, barrier) -> dict:
    """The other kernels of the path on the same games (per-GPU numbers, CUDA events, a few launches each):
    demo generation (metric: demos/s), fused rollout, change of basis, training-sample batcher."""
    import torch

    B = slab.shape[0]
    lay = env.layout(S)
    peak, _ = hbm_peak()
    out = {"per_gpu": True, "n_gpus": world}
    t2, s2 = torch.empty_like(tape3), torch.empty_like(slab)
    ms = _time_ms(lambda: env.make_synthetic_demos(B, R, S, values, probs, shift, seed=1, device=dev, tape=t2, slab=s2), 3, torch)
    algo = S ** 3 + R * 3 * S
    out["demo_gen"] = {"metric": "synthetic_demos_per_sec", "value": B / ms * 1e3, "ms": ms, "R": R,
                       "algorithmic_bytes_per_demo": algo, "hbm_frac": B * algo / (ms * 1e-3) / 1e9 / peak,
                       "bound": "issue slots (Philox draws + packed rank-1 accumulation on the INT pipes), see DESIGN.md 4"}
    ms = _time_ms(lambda: env.accumulate_demos(tape3, S, shift, slab=s2), 3, torch)
    out["demo_accumulate"] = {"value": B / ms * 1e3, "unit": "demos/s", "ms": ms}
    rev = tape3.flip(0).contiguous()
    ms = _time_ms(lambda: env.rollout(slab, rev, S, shift, out=s2), 3, torch)
    algo = 2 * S ** 3 + R * 3 * S + 8
    out["rollout"] = {"metric": "env_steps_per_sec (fused K-step rollout)", "value": B * R / ms * 1e3, "games_per_sec": B / ms * 1e3,
                      "K": R, "ms": ms, "hbm_frac": B * algo / (ms * 1e-3) / 1e9 / peak}
    del rev, t2
    nb = min(B, 1 << 18)
    mats = env.sample_unimodular(nb, S, seed=3, p_nonzero={4: 0.3, 9: 0.08, 16: 0.03}[S], device=dev)
    ms = _time_ms(lambda: env.change_of_basis(slab[:nb], mats, S), 3, torch)
    algo = S ** 3 + 3 * S * S + S ** 3        # what this implementation moves: int8 in, three int8 matrices, int8 out
    algo_survey = S ** 3 + 3 * S * S + 2 * S ** 3  # int8 in, int16 out
    out["change_of_basis"] = {"value": nb / ms * 1e3, "unit": "games/s", "ms": ms, "games": nb,
                              "hbm_frac": nb * algo / (ms * 1e-3) / 1e9 / peak,
                              "hbm_frac_survey_bytes": nb * algo_survey / (ms * 1e-3) / 1e9 / peak,
                              "algorithmic_bytes_per_game": algo, "survey_bytes_per_game": algo_survey,
                              "int_ops_per_game": 6 * S ** 4}
    idx = torch.randint(0, B * R, (1 << 16,), device=dev)
    ms = _time_ms(lambda: env.demo_samples(tape3, slab, idx, S, 2, replay_shift=shift), 3, torch)
    out["demo_sample"] = {"value": idx.numel() / ms * 1e3, "unit": "samples/s", "ms": ms, "dim_t": 2,
                          "hbm_frac": idx.numel() * (2 * S ** 3 * 4) / (ms * 1e-3) / 1e9 / peak}
    # batched leaf expansion (K8): k = 8 candidate actions per state, children + flags + nnz (+ keys)
    nbp = min(B, 1 << 17)
    tape_bk = tape3[:8, :nbp].permute(1, 0, 2).contiguous()
    kk = tape_bk.shape[1]
    moved = lay.game_pitch * (1 + 1 / kk) + lay.token_pitch + 5
    ms = _time_ms(lambda: env.expand_children(slab[:nbp], tape_bk, S, shift, with_keys=False), 3, torch)
    ms_k = _time_ms(lambda: env.expand_children(slab[:nbp], tape_bk, S, shift, with_keys=True), 3, torch)
    out["expand_children"] = {"value": nbp * kk / ms * 1e3, "unit": "children/s", "k": kk, "ms": ms,
                              "hbm_frac": nbp * kk * moved / (ms * 1e-3) / 1e9 / peak,
                              "with_state_keys": {"value": nbp * kk / ms_k * 1e3, "ms": ms_k}}
    del tape_bk
    if S == 9:
        # BASELINE.json quotes the metric at 4x4x4 too: the same two numbers on 2^22 games of the 2x2 matmul size
        # (reference defaults: coefficients {-1,0,1}, P(0) = 0.7, R = 7, shift = 1)
        S4, R4, B4 = 4, 7, 1 << 22
        t4, s4, _ = env.make_synthetic_demos(B4, R4, S4, (-1, 0, 1), (0.15, 0.7, 0.15), 1, seed=2, device=dev)
        ms = _time_ms(lambda: env.make_synthetic_demos(B4, R4, S4, (-1, 0, 1), (0.15, 0.7, 0.15), 1, seed=2, device=dev, tape=t4, slab=s4), 3, torch)
        algo4 = S4 ** 3 + R4 * 3 * S4
        o4, f4, n4 = torch.empty_like(s4), torch.empty(B4, dtype=torch.uint8, device=dev), torch.empty(B4, dtype=torch.int32, device=dev)
        ms_step = _time_ms(lambda: env.step_batch(s4, t4[R4 - 1], S4, 1, out=o4, flags=f4, nnz=n4), 10, torch)
        out["size_4x4x4"] = {"games": B4, "synthetic_demos_per_sec": B4 / ms * 1e3, "demo_hbm_frac": B4 * algo4 / (ms * 1e-3) / 1e9 / peak,
                             "env_steps_per_sec": B4 / ms_step * 1e3,
                             "step_hbm_frac": B4 * ALGO_BYTES[4] / (ms_step * 1e-3) / 1e9 / peak}
        del t4, s4, o4
    if S == 9:
        # BASELINE.json configs[2]: 16x16x16 demos (rank <= 49) followed by the change-of-basis augmentation, one
        # (A, B, C) triple per demo -- the one contraction that runs on the tensor cores (csrc/tg_basis_mma.cu)
        S16, R16, B16 = 16, 49, 1 << 17
        t16, s16, _ = env.make_synthetic_demos(B16, R16, S16, values, probs, shift, seed=3, device=dev)
        ms = _time_ms(lambda: env.make_synthetic_demos(B16, R16, S16, values, probs, shift, seed=3, device=dev, tape=t16, slab=s16), 3, torch)
        algo16 = S16 ** 3 + R16 * 3 * S16
        m16 = env.sample_unimodular(B16, S16, seed=3, p_nonzero=0.03, device=dev)
        ms_cb = _time_ms(lambda: env.change_of_basis(s16, m16, S16), 5, torch)
        ms_cbf = _time_ms(lambda: env.change_of_basis(s16, m16, S16, tape=t16, shift=shift, shift_out=100), 3, torch)
        moved16 = 2 * S16 ** 3 + 3 * S16 * S16
        out["size_16x16x16"] = {"games": B16, "R": R16, "synthetic_demos_per_sec": B16 / ms * 1e3,
                                "demo_hbm_frac": B16 * algo16 / (ms * 1e-3) / 1e9 / peak,
                                "change_of_basis_games_per_sec": B16 / ms_cb * 1e3, "change_of_basis_ms": ms_cb,
                                "change_of_basis_hbm_frac": B16 * moved16 / (ms_cb * 1e-3) / 1e9 / peak,
                                "change_of_basis_hbm_frac_survey_bytes": B16 * (moved16 + S16 ** 3) / (ms_cb * 1e-3) / 1e9 / peak,
                                "change_of_basis_with_factors_games_per_sec": B16 / ms_cbf * 1e3,
                                "kernel": "basis_mma16_kernel (mma.sync int8 + f16, one warp per game)"}
        del t16, s16, m16
    if world > 1:  # demo all-gather timed as its own phase (NVLink-bound)
        from mat_mul_b200 import dist as tgd

        n_g = min(B, 1 << 18)
        shard = slab[:n_g].contiguous()
        barrier()
        ms = _time_ms(lambda: tgd.gather_shards(shard, n_g * world, dim=0), 3, torch)
        out["demo_all_gather"] = {"ms": ms, "bytes_out_per_rank": n_g * world * lay.game_pitch,
                                  "algbw_gbs": n_g * world * lay.game_pitch / (ms * 1e-3) / 1e9}
    return out


def cpu_reference_leg(S: int, shift: int, seconds: float = 12.0):
    """The reference's CPU path for the same workload: the oracle's C port of
    training.py:253-266 on the reference's dtypes (float32 residuals, int64
    tokens), threaded over games with every host core.  Bounded sample."""
    import numpy as np

    from oracle import tg_oracle as orc

    cores = orc.use_all_threads()  # torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm uses every core regardless
    Bs = 1 << 16
    rng = np.random.default_rng(0)
    T = (rng.integers(-2, 3, (Bs, S, S, S)) * (rng.random((Bs, S, S, S)) < 0.3)).astype(np.float32)
    tok = rng.integers(0, 2 * shift + 1, (Bs, 3 * S)).astype(np.int64)
    tok[rng.random((Bs, 3 * S)) < 0.6] = shift
    out = np.empty_like(T); flags = np.empty(Bs, np.uint8); nnz = np.empty(Bs, np.int32)
    orc.step_batch_f32(T, tok, shift, out, flags, nnz)  # warm
    t0 = time.perf_counter()
    orc.step_batch_f32(T, tok, shift, out, flags, nnz)
    one = time.perf_counter() - t0
    reps = max(1, min(20000, int(seconds / max(one, 1e-6))))
    return cores, Bs, reps, T, tok, out, flags, nnz


def run_reference(args) -> None:
    """--impl reference: rank 0 times the CPU port; other ranks exit."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import numpy as np  # noqa: F401

    from oracle import tg_oracle as orc

    S, shift = args.size, 2
    cores, Bs, reps, T, tok, out, flags, nnz = cpu_reference_leg(S, shift, seconds=3.0)
    per_step_reps = max(1, reps // 4)  # one bench "step" = per_step_reps passes over the 65536-game sample
    for _ in range(args.warmup):
        orc.step_batch_f32(T, tok, shift, out, flags, nnz)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for _ in range(per_step_reps):
            orc.step_batch_f32(T, tok, shift, out, flags, nnz)
    dt = time.perf_counter() - t0
    value = args.steps * per_step_reps * Bs / dt
    sample = f"{per_step_reps} passes over {Bs} games of {S}x{S}x{S} per step (float32 residuals, int64 tokens), OpenMP over games"
    line = {
        "impl": "reference", "metric": "env_steps_per_sec", "value": value, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"tensorgame step {S}x{S}x{S}, coefficients -2..2 (CPU port of training.py:253-266)",
                   "games_per_gpu": args.games, "size": S},
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main() -> None:
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    from mat_mul_b200 import _lib, env

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cores = env.bind_host_to_gpu(local)  # pinned staging buffers of the e2e leg land on the GPU's NUMA node
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own log (its version banner at NCCL_DEBUG=VERSION/WARN, which the GPU
        # boxes export) goes to stderr instead
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # honoured above the VERSION level
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":   # the banner-only level always prints to stdout
            del os.environ["NCCL_DEBUG"]
        dist.init_process_group("nccl", device_id=dev)

    S, shift, B, K, W = args.size, 2, args.games, args.steps, max(args.warmup, 3)
    lay = env.layout(S)
    if args.ctas_per_sm:
        _lib.lib().tg_tune_step_ctas_per_sm(args.ctas_per_sm)
    if args.variant:
        _lib.lib().tg_tune_step_variant(args.variant)

    # ---- synthetic games, resident in HBM: the product's own demo generator (Philox stream keyed by the GLOBAL
    # game index, so the union over ranks is independent of N); every game replays its own demo in reverse order
    # (datasets.py:90-92), i.e. a genuine transition on every step; P(coef = 0) = 0.7 as in the reference
    VALUES, PROBS, R = (-2, -1, 0, 1, 2), (0.05, 0.10, 0.70, 0.10, 0.05), {4: 7, 9: 23, 16: 49}[S]
    tape3, slab_a, dflags = env.make_synthetic_demos(B, R, S, VALUES, PROBS, shift, seed=0x5EED, first_demo=rank * B, device=dev)
    slab_b = torch.empty_like(slab_a)
    flags = torch.empty(B, dtype=torch.uint8, device=dev)
    nnz = torch.empty(B, dtype=torch.int32, device=dev)

    def one_step(i: int) -> None:
        src, dst = (slab_a, slab_b) if i % 2 == 0 else (slab_b, slab_a)
        env.step_batch(src, tape3[(R - 1 - i) % R], S, shift, out=dst, flags=flags, nnz=nnz)

    def barrier() -> None:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        one_step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(K):
        one_step(W + i)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    end_flags, end_nnz = flags.clone(), nnz.clone()  # state of the games right after the timed steps
    # keep the clock sampler under the same load for ~1 s (its period is 200 ms); results are discarded
    t_end = time.perf_counter() + max(0.0, 1.0 - ms / 1e3)
    i = 0
    scratch_f, scratch_n = torch.empty_like(flags), torch.empty_like(nnz)
    while time.perf_counter() < t_end:
        env.step_batch(slab_a, tape3[i % R], S, shift, out=slab_b, flags=scratch_f, nnz=scratch_n); i += 1
        if i % 8 == 0:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * K / (ms / 1e3)

    # ---- end to end through the C ABI with HOST buffers (tg_step_host): PCIe in + kernel + PCIe out
    e2e = None
    if not args.no_e2e:
        Be = B
        h_slab = torch.empty((Be, lay.game_pitch), dtype=torch.int8).pin_memory()
        h_slab.copy_(slab_a[:Be])
        h_tape = torch.empty((Be, lay.token_pitch), dtype=torch.uint8).pin_memory()
        h_tape.copy_(tape3[R - 1][:Be])
        h_out = torch.empty_like(h_slab).pin_memory()
        h_flags = torch.empty(Be, dtype=torch.uint8).pin_memory()
        h_nnz = torch.empty(Be, dtype=torch.int32).pin_memory()
        hs = env.HostStepper(S, local, chunk=1 << 16)
        Ke = max(3, min(K, 10))
        for _ in range(2):
            hs.step(h_slab, h_tape, h_out, h_flags, h_nnz, shift)
        barrier()
        t0 = time.perf_counter()
        for _ in range(Ke):
            hs.step(h_slab, h_tape, h_out, h_flags, h_nnz, shift)  # returns with results in host memory
        dt = time.perf_counter() - t0
        hs.close()
        te = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": world * Be * Ke / float(te.item()), "unit": "steps/s",
               "h2d_bytes_per_step": Be * (lay.game_pitch + lay.token_pitch),
               "d2h_bytes_per_step": Be * (lay.game_pitch + 5), "steps": Ke,
               "api": "tg_step_host (C ABI, pinned host buffers, 64Ki-game chunks over 3 streams)",
               "host_cores_bound": len(numa_cores)}
        del h_slab, h_tape, h_out

    # ---- episode statistics: the only collective of the path (two tiny all_reduces), outside the timed region
    from mat_mul_b200 import dist as tgd

    stats = tgd.reduce_episode_stats(end_flags, end_nnz)
    extras = None if args.no_extras else measure_extras(env, dev, S, shift, slab_a, tape3, R, VALUES, PROBS, world, barrier)

    if rank == 0:
        peak, peak_kind = hbm_peak()
        algo = ALGO_BYTES[S] * B
        achieved = algo / (ms / K / 1e3) / 1e9
        traffic = None
        tj = ROOT / "profiles" / "traffic.json"
        if tj.exists():
            try:
                traffic = json.loads(tj.read_text()).get(f"step_S{S}_B{B}")
            except Exception:
                traffic = None
        line = {
            "metric": "env_steps_per_sec", "value": value, "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8",
            "data": "synthetic",
            "config": {"workload": f"tensorgame step {S}x{S}x{S}, 2^{B.bit_length() - 1} games per GPU, coefficients -2..2",
                       "games_per_gpu": B, "size": S, "shift": shift, "sharding": f"game index, {world} rank(s), no data-path collective",
                       "l2": f"inputs larger than L2 ({(2 * lay.game_pitch + lay.token_pitch) * B >> 20} MiB touched per step)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_kind": peak_kind, "kernel": "tg::step_kernel",
                         "algorithmic_bytes_per_step": ALGO_BYTES[S],
                         "moved_bytes_per_step": 2 * lay.game_pitch + lay.token_pitch + 5},
            "e2e": e2e, "gpu_launches": K * world, "clocks": clocks,
            "episode_stats": {"games": stats.games, "solved": stats.solved, "min_nnz": stats.min_nnz,
                              "out_of_range": stats.out_of_range,
                              "note": f"after warmup+steps = {W + K} of the demos' {R} actions, replayed in reverse", "collective": "all_reduce(SUM), all_reduce(MIN) of 5 int64"},
            "extras": extras,
        }
        if world == 1 and not args.no_cpu:
            from oracle import tg_oracle as orc  # CPU baseline leg only (checker never on the product path)

            cores, Bs, reps, T, tok, out, fl, nz = cpu_reference_leg(S, shift)
            t0 = time.perf_counter()
            for _ in range(reps):
                orc.step_batch_f32(T, tok, shift, out, fl, nz)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": reps * Bs / dt, "unit": "steps/s", "cores": cores, "kind": "port",
                                    "sample": f"{reps} passes over {Bs} games ({dt:.1f} s), C port of training.py:253-266 on float32/int64, OpenMP over games"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
