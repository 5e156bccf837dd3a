#!/usr/bin/env python
"""bench.py — train rays/s (fwd+bwd) of the fine-stage render step on B200 (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference render path on host CPU cores

Workload: VoxurfF, 256^3 grids, sparse 100^3 MaskCache, 2^16 rays per GPU per
step, synthetic scene and rays (esr_nerf_b200/synthetic.py), random-init MLPs.  One step =
renderer(**batch) + fixed scalar loss + backward to every parameter .grad (+ one NCCL allreduce of the
gradients when N > 1).  Optimizer / TV / logging excluded.

One JSON line on rank 0; keys per the driver contract (+ roofline, cpu_baseline, e2e, clocks, gpu_launches).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

# stdout carries exactly one JSON line.  Libraries write to file descriptor 1 behind Python's back (NCCL prints its
# version banner there): keep a private handle on the real stdout for the JSON line and point fd 1 at stderr for
# everything else in this process and its children.
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
JSON_OUT = sys.stdout


def _claim_stdout():
    """(bench.py run as a program only — scripts that import this module keep their stdout)"""
    global JSON_OUT
    sys.stdout.flush()
    JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

import torch  # noqa: E402

METRIC = "train rays/sec (fwd+bwd)"
UNIT = "rays/s"
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}

# algorithmic FLOPs per shaded sample, forward: radiance net 85->192->192->192->3, tonemapper 33->192->3
FLOP_RADIANCE = 181_248
FLOP_TONEMAP = 13_824
# tensor-core FLOPs the x2 forward chain EXECUTES per row: three fp16 MMAs per product over the padded 96 -> 192 -> 192 -> 192
# hidden layers (its 192 -> n_out output layer runs on the CUDA cores); reported beside the algorithmic figure
FLOP_X2_EXECUTED = 3 * 2 * (96 * 192 + 2 * 192 * 192)
# what each tensor-core mode delivers against the reference's fp32 nets (tests/test_gpu_voxurff.py, test_gpu_esrnerf.py)
MODE_PARITY = {
    "x2": "outputs ~1e-6, every parameter gradient within 1e-2 of the reference (measured <= 2.5e-3 max-norm, <= 1e-3 rel-L2): "
          "meets north_star's tolerance -> the headline mode",
    "bf16": "outputs 1e-2; MLP / colour-grid gradients 2-5 % rel-L2 (ReLU masks of bf16 pre-activations): does NOT meet "
            "north_star's 1e-2 on gradients -> reported for comparison only",
    "torch_fp32": "library fp32 GEMMs (1e-4 class), not a hand-written path",
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--stage", default="fine", choices=["fine", "lts", "eval"],
                    help="fine: BASELINE configs[1] (default, the bench line); lts: configs[2] shape (ESRNeRF LTS+PDRA "
                         "train step); eval: configs[3] shape (VoxurfF full-image inference, no backward)")
    ap.add_argument("--rays", type=int, default=None, help="rays per GPU per step (default: 2^16 fine, 2^15 lts, "
                                                          "1600x1200/N eval)")
    ap.add_argument("--ltspts", type=int, default=None, help="lts: LTS points per GPU per step (default rays*100/8192)")
    ap.add_argument("--secondary", type=int, default=256, help="lts: secondary rays per LTS point (cfg/app/lts.yaml:41)")
    ap.add_argument("--lts-sampler", default="device", choices=["device", "numpy"],
                    help="lts: draw the LTS points with torch.randperm on the GPU (default) or np.random.choice on the "
                         "host as the reference does (esrnerf.py:792: O(M3) host work, ~5 ms at this size)")
    ap.add_argument("--eval-chunk", type=int, default=1 << 18, help="eval: rays per forward_evaluate call")
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--mask-res", type=int, default=100)
    ap.add_argument("--dense", action="store_true", help="dense MaskCache instead of the sparse shell")
    ap.add_argument("--s-val", type=float, default=20.0)
    ap.add_argument("--mlp-mode", default="x2", choices=["x2", "bf16", "torch_fp32"],
                    help="x2 (default, the headline): tcgen05 chains whose parameter gradients meet north_star's 1e-2 against "
                         "the reference's fp32 nets; bf16: the fast chains (outputs 1e-2, MLP gradients 2-5 %%); torch_fp32: library GEMMs")
    ap.add_argument("--cpu-rays", type=int, default=4096, help="rays per CPU-baseline step (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--profiler-range", action="store_true",
                    help="bracket the timed steps with cudaProfilerStart/Stop (ncu --profile-from-start off captures "
                         "exactly the timed region; numbers printed under a profiler are not bench values)")
    ap.add_argument("--with-optimizer", action="store_true",
                    help="also run the fused Adam step (esr_nerf_b200.optimizer, SURVEY.md §8f row 3) inside the timed "
                         "step; NOT part of the metric BASELINE.json names (render step only), off by default")
    ap.add_argument("--dense-allreduce", action="store_true",
                    help="N>1: all-reduce the dense grid gradients instead of the occupancy-compacted voxel set")
    ap.add_argument("--no-gather", action="store_true",
                    help="eval, N>1: leave every rank's slice of the maps where it is (default: gathered on rank 0 inside the "
                         "timed step, dist.gather_maps: profiles/r02_bench_lines/eval_*_n{2,8}.json)")
    ap.add_argument("--block-exchange", action="store_true",
                    help="N>1: exact two-level exchange of the grid gradients (dist.TouchedBlockCompactor: OR-reduced map of "
                         "touched 8^3 blocks, then pack / all-reduce / unpack of their voxels) instead of the static "
                         "occupancy set; the lts stage's default (profiles/r02_bench_lines/lts_blocks_n*.json)")
    ap.add_argument("--global-rays", type=int, default=None,
                    help="strong scaling: rays per step over ALL GPUs (rays per GPU = global / N; the line says scaling: strong)")
    a = ap.parse_args()
    if a.global_rays is not None:
        a.rays = a.global_rays // max(a.gpus, 1)
    if a.rays is None:
        a.rays = {"fine": 1 << 16, "lts": 1 << 15, "eval": 1600 * 1200 // max(a.gpus, 1)}[a.stage]
    if a.ltspts is None:
        a.ltspts = max(1, a.rays * 100 // 8192)
    if a.stage != "fine" and a.s_val == 20.0:
        a.s_val = 220.0          # lts.yaml:52 / a converged fine-stage model
    return a


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        try:
            d = json.load(open(path))
            out = dict(FALLBACK_PEAKS)
            for k in out:
                if k in d and d[k]:
                    out[k] = float(d[k])
            return out, "measured"
        except Exception:
            pass
    return dict(FALLBACK_PEAKS), "fallback"


def workload_name(a):
    if a.stage == "lts":
        return (f"giftbox_w-like LTS+PDRA stage (ESRNeRF {a.grid}^3, {'dense' if a.dense else 'sparse'} {a.mask_res}^3 "
                f"MaskCache, s_val {a.s_val:g}, pdra_mode), {a.rays} rays + {a.ltspts} LTS points x {a.secondary} "
                f"secondary rays /GPU/step, LTS points drawn on the {a.lts_sampler}, synthetic scene")
    if a.stage == "eval":
        return (f"DTU-like full-image inference (VoxurfF.forward_evaluate {a.grid}^3, {'dense' if a.dense else 'sparse'} "
                f"{a.mask_res}^3 MaskCache, s_val {a.s_val:g}), {a.rays} rays/GPU/image in chunks of {a.eval_chunk}, 12 maps, "
                f"no backward, synthetic scene")
    return (f"giftbox_w-like fine stage (VoxurfF {a.grid}^3, {'dense' if a.dense else 'sparse'} {a.mask_res}^3 MaskCache, "
            f"s_val {a.s_val:g}), {a.rays} rays/GPU/step, synthetic scene")


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port (the reference's Python cannot be imported on the GPU box)
# ---------------------------------------------------------------------------------------------------
def cpu_port_setup(a):
    import esr_testlib as C
    from esr_nerf_b200 import synthetic as S

    torch.set_num_threads(os.cpu_count() or 1)
    scene = C.oracle_scene(a.grid ** 3, a.mask_res, not a.dense)
    torch.manual_seed(0)
    weights = random_mlp_weights()
    params, leaves = C.oracle_params(scene, weights)
    rays = S.make_rays(a.cpu_rays, 1234)
    return scene, params, leaves, rays


def random_mlp_weights():
    """PyTorch-default-initialised MLP weights under the reference's state_dict keys (seed 0)."""
    from esr_nerf_b200.modules import RadianceNet, TonemapNet

    torch.manual_seed(0)
    nets = {"off_rgbnet": RadianceNet(85, 192, 4), "emo_rgbnet": RadianceNet(85, 192, 4), "tonemapper": TonemapNet(33, 192, 2)}
    sd = {}
    for name, net in nets.items():
        for k, v in net.state_dict().items():
            sd[f"{name}.{k}"] = v.detach().clone()
    return sd


def loss_fn(out, rgbs):
    """fixed scalar loss: sRGB MSE + linear MSE + a transmittance term (shape of fine.py:355-382)"""
    l = ((out["srgb/rgb"] + out["etc/white_bg"] - rgbs) ** 2).mean()
    l = l + 0.1 * ((out["lin/rgb"] - rgbs) ** 2).mean()
    l = l + 0.01 * (out["etc/alphainv_cum"] ** 2).mean()
    return l


def cpu_port_step(a, scene, params, leaves, rays):
    from oracle import voxurf_port as P

    for leaf in leaves.values():
        leaf.grad = None
    out, inter = P.voxurff_forward_training(scene, params, rays["rays_o"], rays["rays_d"], rays["viewdirs"],
                                            rays["em_modes"], a.s_val)
    loss_fn(out, rays["rgbs"]).backward()
    return inter


def cpu_coarse_config1(steps: int = 3):
    """BASELINE configs[0] / SURVEY.md §8d "Config 1": the COARSE stage (VoxurfC 64^3, 4096 rays, ~128 candidate samples per
    ray, 57 -> 128 -> 128 -> 3 colour nets, s_val 5), one fwd+bwd on the host cores through the oracle port of the reference's
    VoxurfC render path; 1 warm-up + `steps` timed steps -> rays/s (reported beside the fine-stage CPU number, which is the
    one on the bench line's own workload)"""
    import esr_testlib as C
    from esr_nerf_b200 import synthetic as S
    from oracle import voxurfc_port as PC

    torch.set_num_threads(os.cpu_count() or 1)
    n = 4096
    _, weights = C.load_coarse_case("coarse_sparse_s5")
    scene = C.coarse_oracle_scene(64 ** 3, 32, True)
    params, leaves = C.coarse_oracle_params(scene, weights)
    rays = S.make_rays(n, 2718)
    cot = C.coarse_cotangents(n)

    def step():
        for leaf in leaves.values():
            leaf.grad = None
        out, _ = PC.voxurfc_forward_training(scene, params, rays["rays_o"], rays["rays_d"], rays["viewdirs"], rays["em_modes"], 5.0)
        sum((out[k] * cot[k]).sum() for k in cot).backward()

    step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return {"value": n / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "cores": os.cpu_count() or 1, "kind": "port",
            "workload": "BASELINE configs[0]: coarse stage, VoxurfC 64^3, 4096 rays x ~128 samples, fwd+bwd (oracle port, torch CPU fp32)"}


def run_reference(a, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    scene, params, leaves, rays = cpu_port_setup(a)
    for _ in range(max(a.warmup, 0) and 1):  # one warm-up pass is enough for a CPU port (page-in, thread pool)
        cpu_port_step(a, scene, params, leaves, rays)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        cpu_port_step(a, scene, params, leaves, rays)
    dt = time.perf_counter() - t0
    v = a.cpu_rays * a.steps / dt
    sample = (f"{a.cpu_rays} rays/step of the same workload — the FINE stage of BASELINE configs[1], not the coarse configs[0] "
              f"case — (oracle port of the reference render path, torch CPU, fp32, {cores} threads)")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    try:      # the reference's own CPU-runnable case beside it (an extra: it must never cost the line)
        line["cpu_baseline"]["config1_coarse"] = cpu_coarse_config1()
    except Exception as e:   # noqa: BLE001
        line["cpu_baseline"]["config1_coarse"] = {"error": repr(e)}
    print(json.dumps(line), file=JSON_OUT, flush=True)


# ---------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------
class NvmlClockSampler:
    """SM clock + clocks-event (throttle) reasons of one GPU through NVML (the counters `nvidia-smi
    --query-gpu=clocks.sm,clocks_event_reasons.*` prints), sampled DURING the timed region from the timing thread itself:
    one sample after each step has been queued, while the device is still executing it and the host has milliseconds of
    slack.  A concurrent sampler — the `nvidia-smi -lms` loop of the recipe, or an NVML thread in this process — was
    measured to stall this process at random points (the device-resident region read 11 / 17 / 21 ms per step on runs
    whose per-kernel times and end-to-end region were identical): its driver calls contend with the launches that
    follow a stream-size read, exactly where the host is on the critical path."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}

    def __init__(self, cuda_index):
        import pynvml as N

        N.nvmlInit()
        self.N = N
        pr = torch.cuda.get_device_properties(cuda_index)
        try:
            bus = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
            self.h = N.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            self.h = N.nvmlDeviceGetHandleByIndex(cuda_index)
        self.max_mhz = float(N.nvmlDeviceGetMaxClockInfo(self.h, N.NVML_CLOCK_SM))
        self.reasons_fn = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
        self.sm, self.mask, self.spent = [], 0, 0.0
        self.sample()

    def sample(self):
        t0 = time.perf_counter()
        try:
            self.sm.append(float(self.N.nvmlDeviceGetClockInfo(self.h, self.N.NVML_CLOCK_SM)))
            self.mask |= int(self.reasons_fn(self.h))
        except Exception:
            pass
        self.spent += time.perf_counter() - t0

    def mark(self):
        """forget the samples taken so far (set-up / warm-up): what follows is the timed region"""
        self.sm, self.mask, self.spent = [], 0, 0.0

    def wait_ready(self, timeout=0.0):
        return

    def stop(self):
        sm = sorted(self.sm)
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(v for k, v in self.REASONS.items() if self.mask & k), "samples": len(sm),
               "source": "nvml, one sample per queued step", "host_ms_in_sampler": round(self.spent * 1e3, 3)}
        try:
            self.N.nvmlShutdown()
        except Exception:
            pass
        return out


def make_clock_sampler(cuda_index):
    if os.environ.get("ESR_BENCH_NO_CLOCKS"):      # A/B switch: is the sampler itself visible in the timing?
        return None
    try:
        return NvmlClockSampler(cuda_index)
    except Exception:
        return ClockSampler(cuda_index)     # nvidia-smi loop (B200_PROFILING.md recipe) when NVML is not importable


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def mark(self):
        return

    def sample(self):
        return

    def wait_ready(self, timeout=8.0):
        """block until nvidia-smi has printed its first sample: its start-up (NVML attach, ~1 s) must not land inside
        the timed region, where it stalls the device for tens of milliseconds"""
        t0 = time.time()
        while self.p is not None and self.p.poll() is None and time.time() - t0 < timeout:
            if os.path.getsize(self.f.name) > 0:
                return
            time.sleep(0.05)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            sm.sort()
            out["sm_mhz"] = sm[len(sm) // 2]
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def stage_report(L):
    need = L.esr_stage_timing_report(None, 0)
    buf = ctypes.create_string_buffer(int(need) + 16)
    L.esr_stage_timing_report(buf, len(buf))
    stages = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.split()
        stages[name] = (int(cnt), float(ms))
    return stages


def algorithmic_work(stage, c):
    """(bound, algorithmic bytes or FLOPs of ALL launches of `stage` in one step) — DESIGN.md §4.
    c: per-step counts N, M0, M1, M3, M3_on (rows the emission net runs on)."""
    N, M0, M1, M3, M3on = c["N"], c["M0"], c["M1"], c["M3"], c["M3_on"]
    Mcand = c["Mraw"]
    taps_fwd = (768 + 192) * M3 + 192 * M3on            # 24 SDF taps + off colour (+ emo colour on its rows)
    t = {
        "k_march_count": ("hbm", 24 * N + 32 * M0 + 12 * N),
        "k_march_fill": ("hbm", 24 * N + 32 * M1 + 12 * M1),    # the mask tap (32 M0) is counted once, in the count pass
        "k_neus_alpha": ("hbm", 4 * M1 + 8 * M1),                # read sdf, write alpha + T preset
        "k_transmittance": ("hbm", 4 * M1 + 4 * M1 + 16 * N),     # read alpha, write T; per-ray offsets / count / last
        "k_shade_compact": ("hbm", 8 * M1 + 8 * M3 + 20 * M3),    # read alpha + T (+ step, sdf of survivors), write M3 stream
        # taps at face value + the stream reload; the feature rows / their cotangents (intermediates a fully
        # fused implementation need not materialise) are NOT counted, although these kernels do move them
        "k_encode_fwd": ("hbm", 8 * M3 + taps_fwd),
        "k_encode_bwd": ("hbm", 8 * M3 + 2 * taps_fwd),
        "k_alpha_scan_bwd": ("hbm", 16 * M1 + 8 * M1),
        "k_sdf_scatter": ("hbm", 16 * M1 + 64 * M1),
        "k_composite_fwd": ("hbm", 28 * M3 + 24 * N),
        "k_composite_bwd": ("hbm", 32 * M3 + 28 * M3),
        # forward: off net on every shaded row, emo net on the emission-on rows; backward: each row through ONE net
        # (emission-on rows reach the off net through a stop-gradient, voxurff.py:243-254)
        "k_mlp_fwd_tc_radiance": ("tensor", FLOP_RADIANCE * (M3 + M3on)),
        "k_mlp_fwd_x2_radiance": ("tensor", FLOP_RADIANCE * (M3 + M3on)),
        "k_mlp_dgrad_tc_radiance": ("tensor", FLOP_RADIANCE * M3),
        "k_mlp_wgrad_tc": ("tensor", (FLOP_RADIANCE - 2 * 192 * 3) * M3 + 2 * 33 * 192 * M3),
        "k_mlp_wgrad_tc_out": ("tensor", 2 * 192 * 3 * 2 * M3),
        "k_mlp_fwd_tc_tonemap": ("tensor", FLOP_TONEMAP * M3),
        "k_mlp_dgrad_tc_tonemap": ("tensor", FLOP_TONEMAP * M3),
        # the fused tone-map net: forward; backward = data gradient + weight gradients (its recomputed forward is not
        # algorithmic work)
        "k_tonemap_fwd_fused": ("tensor", FLOP_TONEMAP * M3),
        "k_tonemap_bwd_fused": ("tensor", 2 * FLOP_TONEMAP * M3),
    }
    _ = Mcand
    if "encode_rows" in c:   # lts stage: rows summed over primary / LTS-point / secondary / eps passes (fused.STATS)
        E, Rf, Rb = c["encode_rows"], c["mlp_fwd_rows"], c["mlp_bwd_rows"]
        t.update({
            "k_encode_fwd": ("hbm", (8 + 768 + 384) * E),
            "k_encode_bwd": ("hbm", (8 + 2 * (768 + 384)) * c["encode_bwd_rows"]),
            "k_mlp_fwd_tc_radiance": ("tensor", FLOP_RADIANCE * Rf),
            "k_mlp_fwd_x2_radiance": ("tensor", FLOP_RADIANCE * Rf),
            "k_mlp_dgrad_tc_radiance": ("tensor", FLOP_RADIANCE * Rb),
            "k_mlp_wgrad_tc": ("tensor", FLOP_RADIANCE * Rb + 2 * 33 * 192 * M3),
        })
    return t.get(stage)


def step_roofline(stage_rows, pk):
    """The step's kernels in aggregate, per bound: algorithmic work of all HBM-bound (tensor-bound) launches of one step
    over the time they took together, against the same peaks as `roofline`.  `other_ms_per_step`: library kernels with
    no algorithmic figure (fills, casts, packing)."""
    out = {}
    for bound, unit, peak, scale in (("hbm", "GB/s", pk["hbm_gbs"], 1e9), ("tensor", "TFLOP/s", pk["bf16_tflops_sustained"], 1e12)):
        rows = [r for r in stage_rows if r.get("bound") == bound and r["ms_per_step"] > 0]
        ms = sum(r["ms_per_step"] for r in rows)
        work = sum(r["work_per_step"] for r in rows)
        ach = work / (ms * 1e-3) / scale if ms > 0 else 0.0
        out[bound] = {"kernels": len(rows), "ms_per_step": ms, "algorithmic_work_per_step": work / scale,
                      "work_unit": "GB" if bound == "hbm" else "TFLOP", "achieved": ach, "peak": peak, "unit": unit,
                      "frac": ach / peak if peak else None}
    out["other_ms_per_step"] = sum(r["ms_per_step"] for r in stage_rows if "bound" not in r)
    return out


def lts_loss_fn(out, rgbs):
    """fixed scalar loss with the shape of lts.py:340-420 / pdra.py:390-470: photometric terms + LTS consistency (MSE,
    both sides carry gradient) + normal / emission / BRDF smoothness (L1 against the eps-jittered evaluations) +
    certain-ray emission penalty"""
    l = loss_fn(out, rgbs)
    l = l + ((out["lin/pbr/off"] - out["lin/pbr/off_hat"]) ** 2).mean() + ((out["lin/pbr/emo"] - out["lin/pbr/emo_hat"]) ** 2).mean()
    l = l + 0.1 * (out["etc/normal"] - out["etc/normal_eps"]).abs().mean()
    l = l + 0.1 * (out["etc/emit"] - out["etc/emit_eps"]).abs().mean() + 0.1 * (out["etc/brdf"] - out["etc/brdf_eps"]).abs().mean()
    return l + (out["etc/emit_cert"] ** 2).mean()


def build_stage(a, dev, rank):
    """(model, host batch (pinned), forward kwargs, loss fn or None) for --stage"""
    from esr_nerf_b200 import synthetic as S

    dens = S.mask_density(a.mask_res, not a.dense)
    geo = (S.NEAR, S.FAR, S.BBOX_MIN, S.BBOX_MAX, S.BBOX_MIN, S.BBOX_MAX, S.MASK_ALPHA_INIT, dens, a.s_val, a.grid ** 3)
    host = S.make_rays(a.rays, 1234 + rank)
    if a.stage == "lts":
        from esr_nerf_b200.esrnerf import ESRNeRF

        torch.manual_seed(0)
        model = ESRNeRF(S.lts_cfg(device=str(dev), num_2ndrays=a.secondary, num_ltspts=a.ltspts), *geo)
        model.load_state_dict({**model.state_dict(), **random_mlp_weights()})
        S.fill_esrnerf_model(model)
        model.pdra_mode = True
        model.mlp_mode = a.mlp_mode
        model.lts_sampler = a.lts_sampler
        host["uncert_masks"] = S.uncert_masks(a.rays)
        return model, host, dict(s_val=a.s_val, normal_eps=0.01, emit_eps=0.01), lts_loss_fn
    from esr_nerf_b200.voxurff import VoxurfF

    model = VoxurfF(S.fine_cfg(device=str(dev)), *geo)
    model.load_state_dict({**model.state_dict(), **random_mlp_weights()})
    S.fill_fine_model(model)
    model.mlp_mode = a.mlp_mode
    if a.stage == "eval":
        model.eval()
        host["em_modes"] = torch.zeros_like(host["em_modes"])      # DTU: all emission-off (data/dtu.py:180-183)
        return model, host, dict(pos_rt=torch.eye(3)), None
    return model, host, dict(s_val=a.s_val), loss_fn


def run_b200(a, rank, world, local_rank):
    from esr_nerf_b200 import _lib, fused
    from esr_nerf_b200.dist import GridGradCompactor, TouchedBlockCompactor, allreduce_gradients, gather_maps

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 render path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        import datetime
        # a rank that stops making progress fails the run in minutes (NCCL watchdog) instead of hanging it
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=240))
    L = _lib.lib()

    model, host, fwd_kw, stage_loss = build_stage(a, dev, rank)
    model.keep_streams = True
    params = [p for p in model.parameters() if p.requires_grad]
    compactor = GridGradCompactor(model) if (world > 1 and not a.dense_allreduce and a.stage == "fine") else None
    # the LTS / PDRA stage has no static voxel set (eps-jittered and secondary samples): exact two-level exchange of the
    # touched 8^3 blocks by default (the dense alternative is a 1.28 GB all-reduce per step)
    if world > 1 and a.stage != "eval" and (a.block_exchange or (a.stage == "lts" and not a.dense_allreduce)):
        compactor = TouchedBlockCompactor(model)
    # (round 1 kept this opt-in after a 4-GPU run hung; the likely cause — the collective sequence depended on whether a
    # rank's backward reached the hook — was removed in dist._on_color_grads, and round 2 ran N = 4 three times each with
    # and without it, clean, with the exchange self-check below: gpurun_out/multi_n4 -> profiles/r02_bench_lines)
    # the colour volumes' exchange starts inside the backward pass (dist.GridGradCompactor.overlap_color_allreduce) and
    # runs under the weight-gradient GEMMs; ESR_ALLREDUCE_OVERLAP=0 keeps the whole exchange behind backward()
    overlap = (isinstance(compactor, GridGradCompactor) and not isinstance(compactor, TouchedBlockCompactor)
               and os.environ.get("ESR_ALLREDUCE_OVERLAP", "1") not in ("0", ""))
    if overlap:
        compactor.overlap_color_allreduce(True)
    reduced = [0]
    optimizer = None
    if a.with_optimizer and a.stage != "eval":
        from esr_nerf_b200.optimizer import create_optimizer_or_freeze_model

        lrs = dict(off_color=0.1, off_rgbnet=0.003, emo_color=0.1, emo_rgbnet=0.003, sdf=0.0005, tonemapper=0.003,
                   brdf=0.1, brdfnet=0.001, emitnet=0.001, envmap=0.001)             # cfg/app/{fine,lts}.yaml lrs
        optimizer = create_optimizer_or_freeze_model(model, **lrs)

    host = {k: v.pin_memory() for k, v in host.items()}
    batch = {k: v.to(dev) for k, v in host.items()}

    def train_step(b):
        for p in params:
            p.grad = None
        fused.reset_stats()
        out = model(**fwd_kw, **{k: v for k, v in b.items() if k != "rgbs" or a.stage == "fine"})
        loss = stage_loss(out, b["rgbs"])
        loss.backward()
        if dist is not None:  # rays sharded, gradients summed once per step (north_star)
            reduced[0] = compactor.allreduce() if compactor is not None else allreduce_gradients(params)
        if optimizer is not None:
            optimizer.step()
        return out, loss

    def eval_step(b):
        """one full image: contiguous ray chunks through forward_evaluate, all 12 maps kept (BASELINE configs[3])"""
        outs = []
        n = b["rays_o"].shape[0]
        for lo in range(0, n, a.eval_chunk):
            sl = slice(lo, min(lo + a.eval_chunk, n))
            outs.append(model(rays_o=b["rays_o"][sl], rays_d=b["rays_d"][sl], viewdirs=b["viewdirs"][sl],
                              em_modes=torch.tensor(0), **fwd_kw))
        out = {k: torch.cat([o[k] for o in outs], 0) for k in outs[0]}
        if dist is not None and not a.no_gather:     # configs[3]: the image's maps end up on rank 0 (one gather per image)
            full = gather_maps(out, n * world, rank, world)
            out = full if full is not None else out
        return out, out["etc/depth"].mean()

    step = eval_step if a.stage == "eval" else train_step

    def sync_all():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def exchange_check():
        """first untimed step at N > 1: the step's exchange (compacted / block / early-started) against a plain dense
        all-reduce of the same local gradients.  Two ranks: two-term sums, exact in any order -> bit-equal required;
        more ranks: the summation order differs with the message size -> 1e-6 of the tensor's largest magnitude."""
        for p in params:
            p.grad = None
        out = model(**fwd_kw, **{k: v for k, v in batch.items() if k != "rgbs" or a.stage == "fine"})
        stage_loss(out, batch["rgbs"]).backward()
        local = {p: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)) for p in params}
        dense = {p: g.clone() for p, g in local.items()}
        for g in dense.values():
            dist.all_reduce(g, op=dist.ReduceOp.SUM)
        if compactor is not None:
            compactor.allreduce()
        else:
            allreduce_gradients(params)
        worst, bad = 0.0, []
        for name, p in model.named_parameters():
            if p not in dense:
                continue
            got = p.grad if p.grad is not None else torch.zeros_like(p)
            scale = float(dense[p].abs().max())
            err = float((got - dense[p]).abs().max()) / max(scale, 1e-30)
            worst = max(worst, err)
            if err > (0.0 if world == 2 else 1e-6):
                bad.append((name, err))
        t = torch.tensor([worst, float(len(bad))], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res = {"status": "ok" if t[1].item() == 0 else "MISMATCH", "worst_rel_to_max": float(t[0].item()),
               "against": "dense all_reduce(SUM) of the same local gradients, every parameter",
               "tolerance": "bit-equal" if world == 2 else "1e-6 of the tensor's max magnitude (summation order)",
               "exchange": type(compactor).__name__ if compactor is not None else "dense"}
        if bad:
            res["mismatched"] = bad[:8]
        return res

    exchange = exchange_check() if (dist is not None and a.stage != "eval") else None

    # clocks / throttle reasons are sampled from before the warm-up to the end of the timed region (nvidia-smi needs a
    # moment to start; starting it inside the timed region would both miss it and perturb it)
    sampler = make_clock_sampler(local_rank) if rank == 0 else None
    for _ in range(max(a.warmup, 5)):
        step(batch)
    sync_all()
    if sampler:
        sampler.wait_ready()
    st = model.last_streams["streams"]
    counts = {"N": st.n_rays, "Mraw": int(st.n_steps.sum()), "M0": int(st.cnt_inbox.sum()), "M1": st.m1, "M3": st.m3,
              "M3_on": st.m3_on}
    if a.stage == "eval":   # the last chunk only is kept in last_streams: scale to the image
        f = a.rays / max(st.n_rays, 1)
        counts = {k: int(v * f) for k, v in counts.items()}
        counts["M3_on"] = counts["M3"]
    if a.stage == "lts":
        st2 = model.last_streams["lts"]["streams"]
        counts.update(secondary_rays=st2.n_rays, M0_secondary=int(st2.cnt_inbox.sum()), M1_secondary=st2.m1,
                      M3_secondary=st2.m3, encode_rows=fused.STATS["encode_rows"], mlp_fwd_rows=fused.STATS["mlp_fwd_rows"],
                      mlp_bwd_rows=fused.STATS["mlp_bwd_rows"],
                      encode_bwd_rows=fused.STATS["encode_rows"])

    # ---- device-resident timed region (value): exactly K steps between two CUDA events ----
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the host must stay ahead of the device between the two stream-size reads of a step: a cyclic-GC pass over the
    # interpreter's heap (tens of ms with torch loaded) inside the timed region shows up as device idle time, so the
    # collector is run now and kept off while timing (what training loops do with gc.freeze / manual collection)
    import gc
    gc.collect()
    gc.freeze()
    gc.disable()
    # the housekeeping above (stream-count read-backs, garbage collection: ~0.2 s with the device idle) is followed by a
    # transient: the SECOND step after it stalls for 25-215 ms (per-step events, 4 of 8 runs; steps 3..K then sit at
    # 10.90 +- 0.03 ms).  Three more untimed steps absorb it, exactly as the end-to-end region below has always been
    # entered; the timed region still starts from a barrier + synchronize with nothing in flight.
    for _ in range(3):
        step(batch)
    sync_all()
    launches0 = L.esr_launch_count()
    if sampler:
        sampler.mark()          # clock samples from here on belong to the timed region
    if a.profiler_range:
        torch.cuda.profiler.start()
    e0.record()
    marks = []
    for _ in range(a.steps):
        step(batch)
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()             # per-step boundaries (diagnostic: "ms_each"); value is e0 -> e1 over all K steps
        marks.append(ev)
        if sampler:
            sampler.sample()    # the step is queued and still executing: clocks under load, host off the critical path
    e1.record()
    sync_all()
    ms_each = [round(x.elapsed_time(y), 3) for x, y in zip([e0] + marks[:-1], marks)]
    if a.profiler_range:
        torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1)
    launches = L.esr_launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())

    # ---- the same K steps again with a CUDA-event pair around every library launch (roofline table) ----
    L.esr_stage_timing(1)
    for _ in range(a.steps):
        step(batch)
    sync_all()
    stages = stage_report(L)
    L.esr_stage_timing(0)

    # ---- end-to-end region: pinned host rays -> H2D, step, D2H of the rendered outputs ----
    e2e = None
    if not a.no_e2e:
        h2d = sum(v.numel() * v.element_size() for v in host.values())
        d2h = 0

        pinned_out = {}

        def to_host(name, t):
            """device -> pinned host buffer on the current stream (buffers are reused; one synchronisation per step)"""
            t = t.detach()
            buf = pinned_out.get(name)
            if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
                buf = pinned_out[name] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            buf.copy_(t, non_blocking=True)
            return buf

        copy_stream = torch.cuda.Stream(device=dev)

        def stage_in():
            """H2D of one step's inputs from pinned host memory, on a copy stream: issued right after the previous
            step has been queued, it overlaps that step's kernels (a prefetching loader); exactly one per step"""
            with torch.cuda.stream(copy_stream):
                b = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return b, ev

        def e2e_step(staged, stage_next):
            b, ev = staged
            cur = torch.cuda.current_stream()
            cur.wait_event(ev)
            for v in b.values():
                v.record_stream(cur)
            out, loss = step(b)
            nxt = stage_in() if stage_next else None
            keys = sorted(out) if a.stage == "eval" else ["srgb/rgb", "lin/rgb", "etc/alphainv_cum"]
            res = [to_host(k, out[k]) for k in keys] + [to_host("loss", loss)]
            cur.synchronize()                                  # the step's results are on the host from here on
            return sum(r.numel() * r.element_size() for r in res), nxt

        staged = stage_in()
        for _ in range(2):   # untimed: first use of the H2D / D2H staging buffers
            _, staged = e2e_step(staged, True)
        sync_all()
        e0.record()
        staged = stage_in()                                    # K copies inside the region: this one + K - 1 prefetched
        for i in range(a.steps):
            d2h, staged = e2e_step(staged, i + 1 < a.steps)
        e1.record()
        sync_all()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": a.rays * world * a.steps / (float(t.item()) * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "pipeline": "inputs: pinned host -> device on a copy stream, step k+1's copy overlaps step k; results: "
                           "device -> pinned host + one stream synchronisation every step"}

    # ---- the other tensor-core mode, same K steps (reported beside the headline, never as the headline) ----
    other_mode = None
    if a.stage != "eval" and a.mlp_mode in ("x2", "bf16"):
        alt = "bf16" if a.mlp_mode == "x2" else "x2"
        model.mlp_mode = alt
        for _ in range(3):
            step(batch)
        sync_all()
        e0.record()
        for _ in range(a.steps):
            step(batch)
        e1.record()
        sync_all()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        other_mode = {"mlp_mode": alt, "ms_per_step": float(t.item()) / a.steps,
                      "value": a.rays * world * a.steps / (float(t.item()) * 1e-3), "unit": UNIT,
                      "parity": MODE_PARITY[alt]}
        model.mlp_mode = a.mlp_mode

    gc.enable()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    pk, pk_src = peaks()
    stage_rows = []
    for name, (cnt, tot) in stages.items():
        w = algorithmic_work(name, counts)
        row = {"kernel": name, "launches_per_step": cnt / a.steps, "ms_per_step": tot / a.steps}
        if w is not None:
            bound, work = w
            row["work_per_step"] = float(work)
            per_s = work / (tot / a.steps * 1e-3) if tot > 0 else 0.0
            if bound == "hbm":
                row.update(bound="hbm", achieved=per_s / 1e9, unit="GB/s", frac=per_s / 1e9 / pk["hbm_gbs"])
            else:
                row.update(bound="tensor", achieved=per_s / 1e12, unit="TFLOP/s",
                           frac=per_s / 1e12 / pk["bf16_tflops_sustained"])
                if name == "k_mlp_fwd_x2_radiance":   # the fp32-class forward costs three fp16 MMAs per algorithmic product
                    ex = per_s * FLOP_X2_EXECUTED / FLOP_RADIANCE
                    row.update(executed_tflops=ex / 1e12, frac_executed=ex / 1e12 / pk["bf16_tflops_sustained"])
        if name == "k_encode_bwd":   # (a fraction above 1 is this accounting, not bandwidth)
            row["note"] = ("face-value tap bytes (SURVEY 8d: 4 B x channels x 8 corners per tap, twice); the kernel sums the "
                           "contributions of consecutive samples to a voxel in the warp and issues one RED for the run, so "
                           "fewer bytes than that reach the L2 (profiles/r02_encode_bwd_merged.md)")
        stage_rows.append(row)
    stage_rows.sort(key=lambda r: -r["ms_per_step"])
    top = next((r for r in stage_rows if "bound" in r), None)
    roofline = None
    if top is not None:
        traffic = None   # DRAM bytes per launch of this kernel from the committed ncu --set full capture of this workload
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if a.stage == "fine" and a.rays == (1 << 16) and a.grid == 256 and not a.dense and os.path.isfile(tpath):
            traffic = json.load(open(tpath))["dram_bytes_per_launch"].get(top["kernel"])
        roofline = {"kernel": top["kernel"], "bound": top["bound"], "achieved": top["achieved"],
                    "peak": pk["hbm_gbs"] if top["bound"] == "hbm" else pk["bf16_tflops_sustained"],
                    "unit": top["unit"], "frac": top["frac"], "traffic": traffic, "peak_source": pk_src,
                    "avg_launch_ms": top["ms_per_step"] / max(top["launches_per_step"], 1e-9),
                    "share_of_kernel_time": top["ms_per_step"] / max(sum(r["ms_per_step"] for r in stage_rows), 1e-9)}

    try:
        step_rf = step_roofline(stage_rows, pk)
    except Exception as e:   # a reporting extra must never cost the bench line
        step_rf = {"error": repr(e)}

    cpu_baseline = None
    if world == 1 and not a.no_cpu_baseline and a.stage == "fine":
        # the CPU leg runs in its own process (this arm's address space never maps oracle/): bench.py --impl reference,
        # 1 warm-up + 2 timed steps of the bounded sample
        import subprocess
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "1",
               "--cpu-rays", str(a.cpu_rays), "--grid", str(a.grid), "--mask-res", str(a.mask_res), "--s-val", str(a.s_val)]
        if a.dense:
            cmd.append("--dense")
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
            ref_line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
            cpu_baseline = dict(ref_line["cpu_baseline"])
            cpu_baseline["sample"] += ", 1 warm-up + 2 timed steps, separate process"
        except Exception as e:   # noqa: BLE001 — the baseline leg must never cost the bench line
            cpu_baseline = {"error": repr(e)}

    total_rays = a.rays * world * a.steps
    metric = {"fine": METRIC, "lts": "train rays/sec (fwd+bwd), LTS+PDRA stage",
              "eval": "render rays/sec (inference, 12 maps)"}[a.stage]
    line = {
        "metric": metric, "value": total_rays / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": a.steps,
        "warmup": max(a.warmup, 5), "ms_per_step": ms / a.steps, "higher_is_better": True,
        "scaling": "strong" if a.global_rays is not None else "weak",
        "vs_baseline": None,
        "dtype": {"bf16": "f32 (grids, scan, compositing) + bf16 tensor-core MLPs, f32 accumulate",
                  "x2": "f32 (grids, scan, compositing) + tensor-core MLPs on fp16 operands (forward: hi + lo pairs, three "
                        "MMAs per product; backward: scaled fp16), f32 accumulate",
                  "torch_fp32": "f32"}[a.mlp_mode],
        "mlp_mode": a.mlp_mode, "mlp_mode_parity": MODE_PARITY[a.mlp_mode], "other_mode": other_mode,
        "exchange_check": exchange["status"] if exchange else None, "exchange_check_detail": exchange,
        "data": "synthetic",
        "config": {"workload": workload_name(a),
                   "parallelism": f"dp{world} (rays sharded, one gradient allreduce per step"
                                  + (f", {reduced[0] / 1e6:.0f} MB/rank" if world > 1 else "") + ")",
                   "stage": a.stage, "optimizer_in_step": bool(optimizer is not None),
                   "l2": "working set (0.83 GB of grids + grads) exceeds the 126 MB L2; no flush between steps",
                   "host": "python cyclic GC collected + frozen before, disabled during the timed regions",
                   "counts_per_gpu_step": counts,
                   "samples_per_s": {"candidate_M0": counts["M0"] * world * a.steps / (ms * 1e-3),
                                     "shaded_M3": counts["M3"] * world * a.steps / (ms * 1e-3)}},
        "roofline": roofline, "step_roofline": step_rf, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": clocks, "ms_each": ms_each, "kernels": stage_rows[:16],
    }
    print(json.dumps(line), file=JSON_OUT, flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    _claim_stdout()
    a = parse()
    # a wedged device or collective must end the process, not the box's time limit: after ESR_BENCH_WATCHDOG_S seconds
    # (default 20 min + 1 s per step; a default run takes ~2 min) every thread's stack goes to stderr and the process exits non-zero
    import faulthandler
    if a.impl != "reference":    # (the CPU arm cannot wedge, and its run time is whatever --steps asks for)
        faulthandler.dump_traceback_later(float(os.environ.get("ESR_BENCH_WATCHDOG_S", 1200 + a.steps)), exit=True,
                                          file=sys.stderr)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference(a, rank, world)
    else:
        run_b200(a, rank, world, local_rank)


if __name__ == "__main__":
    main()
