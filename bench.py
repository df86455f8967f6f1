#!/usr/bin/env python
"""bench.py -- oscillation-reweighted AND histogrammed events/s (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

Workload (config.workload = "C3"): BASELINE.json configs[2] -- synthetic per-event prob3
reweighting through the 12-layer PREM Earth + weighted histogramming (sumw2) into the
`dragon_datarelease` 8x8x2 binning, 1e8 events per GPU in 12 flavour containers, FP64.
(configs[1], the IceCube-3y pipeline on its bundled MC, cannot be run: the MC file is absent
from the reference tree; configs[0] is the reference's CPU-sized grid case and is a parity test.)

One "step" = one template evaluation = one pass of the hot path over all events of this rank:
for each container the fused kernel (layers -> prob3 -> weights *= flux.prob -> histogram w, w^2),
then ONE exchange of the [12, 2, 128] histogram buffer when N > 1 (weak scaling: events/GPU fixed).
All 12 containers are evaluated by ONE launch of the fused kernel.

value : events/s with the event arrays resident in HBM (inputs 4.4 GB/GPU >> 126 MB L2, so no
        L2 flush is needed between steps).
e2e   : the same step through ReweightEngine.evaluate_host -- event arrays in pinned HOST memory,
        H2D copies of every input and the D2H read of the histograms inside the timed region.
roofline : dominant kernel = reweight_hist_kernel, FP64-compute bound.  achieved =
        algorithmic FLOPs (1329 + 2367 L_distinct + 226 L_cached per event, counted from the
        actual coszen array) / CUDA-event time of the launches.  MEASURED_PEAKS.json has no FP64
        entry, so the peak is a DFMA micro-benchmark run in this process (pisab_fp64_peak_probe);
        the nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz = 37.2 TFLOP/s is printed beside it.
cpu_baseline : the oracle port (oracle/*.c, OpenMP over all host cores) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "osc-weighted events/sec (prob3+hist)"
UNIT = "events/s"
NOMINAL_FP64_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="native", choices=["native", "reference"])
    p.add_argument("--events-per-gpu", type=float, default=1e8)
    p.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    p.add_argument("--nsi", action="store_true", help="standard-NSI matter potential (config C4)")
    p.add_argument("--f32-math", default="mixed", choices=["mixed", "fp64"],
                   help="arithmetic behind --dtype f32: mixed precision (the FP32 mode) or FP64 on float32 storage")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--cpu-seconds", type=float, default=12.0)
    return p.parse_args()


def config_dict(args, n_per_gpu, world):
    return {
        "workload": "C3: synthetic per-event prob3 through PREM_12layer + weighted hist (sumw2), "
                    "dragon_datarelease 8x8x2, 12 flavour containers" + (", standard NSI" if args.nsi else ""),
        "events_per_gpu": int(n_per_gpu), "global_events": int(n_per_gpu) * world, "n_bins": 128,
        "containers": 12, "earth_model": "PREM_12layer", "osc": "nufit v2.0 NH",
        "l2": "inputs (%.1f GB/GPU) larger than L2; no flush" % (n_per_gpu * 44 / 1e9),
        "parallelism": "events sharded over %d GPU(s); one all-reduce of [12,2,128] f64 per step" % world,
    }


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port) on host cores
# ------------------------------------------------------------------------------------------------
class CpuChain:
    """layers -> propagate_array -> fill_probs -> weights *= flux.prob -> histogram (w, w^2), per
    container, exactly the reference's stage sequence (prob3.py:406-409,581-622; hist.py:198-209)."""

    def __init__(self, nsi=False):
        import oracle
        from pisa_b200.utils import synthetic as syn
        self.oracle, self.syn = oracle, syn
        self.threads = os.cpu_count() or 1
        prem = np.loadtxt(os.path.join(ROOT, "pisa_b200", "resources", "osc", "PREM_12layer.dat"))
        self.L = oracle.OracleLayers(prem, syn.EARTH["detector_depth"], syn.EARTH["prop_height"])
        self.L.setElecFrac(syn.EARTH["YeI"], syn.EARTH["YeO"], syn.EARTH["YeM"])
        self.dm, self.mix, self.mat_pot = syn.osc_matrices(nsi=syn.STD_NSI if nsi else None)
        self.zero_c = np.zeros((3, 3), dtype=np.complex128)
        self.zero_f = np.zeros((3, 3))

    def make(self, n, seed=1):
        ev = self.syn.make_events_numpy(n, seed)
        per = n // 12
        self.blocks = []
        for c, (name, nubar, flav) in enumerate(self.syn.CONTAINERS):
            sl = slice(c * per, (c + 1) * per if c < 11 else n)
            self.blocks.append((nubar, flav, {k: v[sl] for k, v in ev.items()}))
        self.n = n

    def step(self):
        o = self.oracle
        out = np.zeros((12, 2, 128))
        for c, (nubar, flav, ev) in enumerate(self.blocks):
            _, den, dis = self.L.calcLayers(ev["true_coszen"])
            prob = o.propagate_array(self.dm, self.mix, self.mat_pot, -1, self.zero_c, self.zero_f, nubar,
                                     ev["true_energy"], den, dis, n_threads=self.threads)
            pe, pmu = o.fill_probs(prob, 0, flav), o.fill_probs(prob, 1, flav)
            w = ev["weights"] * (ev["nu_flux"][:, 0] * pe + ev["nu_flux"][:, 1] * pmu)
            ie = o.digitize_irregular(ev["reco_energy"], self.syn.DRAGON_E_EDGES)
            i2, _ = o.regular_index([ev["reco_coszen"], ev["pid"]], [-1.0, -0.5], [1.0, 1.5], [8, 2])
            idx = np.where((ie >= 0) & (ie < 8) & (i2 >= 0), ie * 16 + i2, -1)
            out[c, 0] = o.accumulate(idx, w, 128)
            out[c, 1] = o.accumulate(idx, w * w, 128)
        return out

    def sized_for(self, seconds):
        """events for ~`seconds` of CPU work per step (pilot run)."""
        self.make(24000)
        self.step()  # loads / warms the library
        t0 = time.perf_counter()
        self.step()
        rate = 24000 / (time.perf_counter() - t0)
        return int(min(max(rate * seconds, 48000), 2e7)) // 12 * 12, rate


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    chain = CpuChain(nsi=args.nsi)
    per_step_s = max(2.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    n, _ = chain.sized_for(per_step_s)
    chain.make(n)
    for _ in range(args.warmup):
        chain.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        chain.step()
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = "%d synthetic events/step (same laws as the C3 workload, bounded sample)" % n
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, int(args.events_per_gpu), world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": chain.threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled through NVML in a thread (every ~5 ms) between start() and
    stop(); the timed region is short (~0.1 s), so `nvidia-smi -lms` (>= 100 ms period, slow start) would
    miss it.  Falls back to one `nvidia-smi` query per start/stop when pynvml is unavailable."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4), ("hw_power_brake_slowdown", 0x80))

    def __init__(self, gpu_index):
        self.gpu, self.sm, self.mask, self.power = gpu_index, [], 0, []
        self.handle, self.nv, self.run, self.thread, self.sm_max = None, None, False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES if it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = gpu_index
            if vis:
                try:
                    phys = int(vis.split(",")[gpu_index])
                except (ValueError, IndexError):
                    phys = gpu_index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001  (any NVML problem -> fallback)
            self.handle = None

    def _sample(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
        try:
            self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        except Exception:  # noqa: BLE001
            try:
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            except Exception:  # noqa: BLE001
                pass
        try:
            self.power.append(nv.nvmlDeviceGetPowerUsage(self.handle) / 1e3)
        except Exception:  # noqa: BLE001
            pass

    def _loop(self):
        while self.run:
            self._sample()
            time.sleep(0.005)

    def _smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=10).stdout.strip().split(",")
            self.sm.append(float(out[0]))
            self.sm_max = float(out[1])
            for (nm, bit), v in zip(self.REASONS, out[2:6]):
                if v.strip().lower().startswith("active"):
                    self.mask |= bit
        except Exception:  # noqa: BLE001
            pass

    def start(self):
        if self.handle is None:
            self._smi()
            return
        self.run = True
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        if self.handle is None:
            self._smi()
        else:
            self._sample()
            self.run = False
            self.thread.join(timeout=1)
        reasons = sorted(nm for nm, bit in self.REASONS if self.mask & bit)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": reasons, "samples": len(self.sm),
                "power_w_max": max(self.power) if self.power else None,
                "how": "NVML polled every 5 ms during the timed region" if self.handle is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------
def run_native(args):
    import torch
    import torch.distributed as dist
    from pisa_b200 import ops
    from pisa_b200.engine import ReweightEngine
    from pisa_b200.stages.osc.layers import Layers
    from pisa_b200.utils import synthetic as syn

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the native arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    dtype = np.float64 if args.dtype == "f64" else np.float32
    ops.set_f32_math(args.f32_math)
    n_gpu = int(args.events_per_gpu) // 12 * 12

    L = Layers(os.path.join(ROOT, "pisa_b200", "resources", syn.EARTH["earth_model"]),
               syn.EARTH["detector_depth"], syn.EARTH["prop_height"])
    L.setElecFrac(syn.EARTH["YeI"], syn.EARTH["YeO"], syn.EARTH["YeM"])
    earth = L.earth_struct()
    dm, mix, mat_pot = syn.osc_matrices(nsi=syn.STD_NSI if args.nsi else None)
    consts = ops.OscConsts.from_matrices(dm, mix, mat_pot)
    binning, keep = ops.make_binning(syn.DRAGON_DIMS, dev)

    eng = ReweightEngine(earth, syn.DRAGON_NBINS, dtype, dev)       # resident arrays
    eng_host = ReweightEngine(earth, syn.DRAGON_NBINS, dtype, dev)  # pinned host arrays (e2e)
    per = n_gpu // 12
    sum_distinct = sum_cached = 0.0
    for c, (name, nubar, flav) in enumerate(syn.CONTAINERS):
        ev = syn.make_events_torch(per, seed=1000 * rank + c + 1, dtype=dtype, device=dev)
        # setup-time work, like hist.setup_function / Container.translate: static bin index
        index = ops.hist_index(binning, [ev["reco_energy"], ev["reco_coszen"], ev["pid"]])
        _, distinct, cached = syn.layer_counts(L.coszen_limit, ev["true_coszen"])
        sum_distinct += float(distinct.sum())
        sum_cached += float(cached.sum())
        arrays = dict(true_energy=ev["true_energy"], true_coszen=ev["true_coszen"], nu_flux=ev["nu_flux"],
                      weights=ev["weights"], index=index)
        eng.add_container(name, nubar, flav, **arrays)
        if not args.no_e2e:
            eng_host.add_container(name, nubar, flav, **{k: v.cpu() for k, v in arrays.items()})
        del ev
    flops_event = syn.flops_per_event(sum_distinct / n_gpu, sum_cached / n_gpu)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # ---- resident-input steps ------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        eng.evaluate(consts)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    ops.launch_count(reset=True)
    kernel_events = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.nvtx.range_push("timed")   # ncu --nvtx --nvtx-include "timed/" lists exactly these launches
    e0.record()
    for _ in range(args.steps):
        out = eng.evaluate(consts, events=kernel_events)
    e1.record()
    barrier()
    torch.cuda.nvtx.range_pop()
    launches = ops.launch_count()
    clocks = sampler.stop()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    value = n_gpu * world / (ms_step * 1e-3)
    # per-launch CUDA-event times bracket the fused kernel (ONE launch for all 12 containers) + its tiny
    # partial-reduction kernel
    launch_ms = float(np.mean([a.elapsed_time(b) for a, b in kernel_events]))
    per = n_gpu  # events per fused launch
    hist_total = float(out[:, 0].sum())

    # ---- FP64 roofline denominator -------------------------------------------------------------
    peak_flops, _ = ops.fp64_peak_probe(20000)
    achieved = flops_event * per / (launch_ms * 1e-3)
    # DRAM traffic per launch from the committed ncu --set full capture (bytes/event measured there
    # on a 1e6-event launch of the same kernel; it scales linearly with the events of a launch)
    traffic = pipe_active = executed_flop = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tr = json.load(f)["reweight_hist_kernel<%s>" % ("double" if args.dtype == "f64" else "float")]
        traffic = tr["dram_bytes_per_event"] * per
        pipe_active = tr.get("fp64_pipe_active_pct")
        executed_flop = tr.get("executed_fp64_flop_per_event")
    except (OSError, KeyError, ValueError):
        pass
    roofline = {
        "bound": "fp64", "kernel": "reweight_hist_kernel<double>", "achieved": achieved / 1e12,
        "peak": peak_flops / 1e12, "unit": "TFLOP/s", "frac": achieved / peak_flops, "traffic": traffic,
        "peak_source": "DFMA micro-benchmark in this process (MEASURED_PEAKS.json has no FP64 entry); "
                       "nominal %.1f TFLOP/s" % NOMINAL_FP64_TFLOPS,
        "algorithmic_flops_per_event": flops_event, "events_per_launch": per, "launch_ms": launch_ms,
        "frac_of_nominal": achieved / (NOMINAL_FP64_TFLOPS * 1e12),
        "algorithmic_bytes_per_event": 44 if args.dtype == "f64" else 24,
        "hbm_gbs": (44 if args.dtype == "f64" else 24) * per / (launch_ms * 1e-3) / 1e9,
        # `frac` follows the contract (reference arithmetic, SURVEY 8d, / measured DFMA peak) and exceeds 1
        # because the kernel's formulation executes ~5x fewer FLOPs than the reference's; the hardware-side
        # figures come from the committed ncu capture of the same kernel (profiles/ncu_traffic.json):
        "fp64_pipe_active_pct_ncu": pipe_active,
        "executed_flops_per_event_ncu": executed_flop,
        "executed_tflops": None if executed_flop is None else executed_flop * per / (launch_ms * 1e-3) / 1e12,
        "executed_frac": None if executed_flop is None else executed_flop * per / (launch_ms * 1e-3) / peak_flops,
    }

    # ---- end to end (host buffers) -------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        for _ in range(2):
            eng_host.evaluate_host(consts)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            host_out = eng_host.evaluate_host(consts)
        torch.cuda.synchronize()
        t_e2e = max_over_ranks(time.perf_counter() - t0)
        barrier()
        # raw pinned-host -> device copy bandwidth of this box, as the yardstick for the e2e number
        probe = torch.empty(1 << 28, dtype=torch.uint8).pin_memory()
        probe_d = torch.empty(1 << 28, dtype=torch.uint8, device=dev)
        probe_d.copy_(probe, non_blocking=True)
        torch.cuda.synchronize()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(4):
            probe_d.copy_(probe, non_blocking=True)
        p1.record()
        torch.cuda.synchronize()
        h2d_peak = 4 * probe.numel() / (p0.elapsed_time(p1) * 1e-3) / 1e9
        del probe, probe_d
        e2e = {"value": n_gpu * world * args.steps / t_e2e, "unit": UNIT,
               "h2d_bytes_per_step": int(eng_host.last_h2d_bytes), "d2h_bytes_per_step": int(eng_host.last_d2h_bytes),
               "ms_per_step": 1e3 * t_e2e / args.steps,
               "h2d_gbs": eng_host.last_h2d_bytes * args.steps / t_e2e / 1e9, "h2d_peak_gbs_measured": h2d_peak,
               "api": "pisa_b200.engine.ReweightEngine.evaluate_host (pinned host arrays, double-buffered H2D)"}
        if world == 1 and abs(float(host_out[:, 0].sum()) / hist_total - 1) > 1e-9:
            raise SystemExit("bench.py: e2e and resident histograms disagree")

    # ---- CPU baseline on rank 0, N = 1 only ----------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        chain = CpuChain(nsi=args.nsi)
        n_cpu, _ = chain.sized_for(args.cpu_seconds)
        chain.make(n_cpu)
        t0 = time.perf_counter()
        chain.step()
        dt = time.perf_counter() - t0
        cpu = {"value": n_cpu / dt, "unit": UNIT, "cores": chain.threads, "kind": "port",
               "sample": "%d synthetic events, one step of the oracle port (OpenMP, %d threads)" % (n_cpu, chain.threads)}

    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "config": config_dict(args, n_gpu, world),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        }
        return line
    return None


class _StdoutToStderr:
    """Route fd 1 to stderr while libraries initialise (NCCL prints its version banner on stdout), so that
    the ONE JSON line is the only thing this program writes to stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        with _StdoutToStderr():
            line = run_native(args)
        if line is not None:
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
