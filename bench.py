#!/usr/bin/env python
"""bench.py -- oscillation-reweighted AND histogrammed events/s (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own numba CPU path (baseline/_ref)

Workload (config.workload = "C3"): BASELINE.json configs[2] -- synthetic per-event prob3
reweighting through the 12-layer PREM Earth + weighted histogramming (sumw2) into the
`dragon_datarelease` 8x8x2 binning, 1e8 events per GPU in 12 flavour containers, FP64.
(configs[1], the IceCube-3y pipeline on its bundled MC, cannot be run: the MC file is absent
from the reference tree; configs[0] is the reference's CPU-sized grid case and is a parity test.)

One "step" = one template evaluation = one pass of the hot path over all events of this rank:
ONE launch of the fused kernel over all 12 containers (layers -> prob3 -> weights *= flux.prob ->
histogram w, w^2), then ONE exchange of the [12, 2, 128] histogram buffer when N > 1 (weak scaling:
events/GPU fixed).

value    : events/s with the event arrays resident in HBM (inputs 4.4 GB/GPU >> 126 MB L2, so no
           L2 flush is needed between steps).
e2e      : the same step through ReweightEngine.evaluate_host -- event arrays in pinned HOST memory,
           H2D copies of EVERY input and the D2H read of the histograms inside the timed region.
           `e2e_changed_only` repeats it with only the arrays a fit changes (weights, nu_flux) travelling.
roofline : dominant kernel = reweight_hist_kernel, FP64-pipe bound.  MEASURED_PEAKS.json has no
           FP64 entry, so `peak` is a DFMA micro-benchmark run in this process (pisab_fp64_peak_probe).
           `frac` = EXECUTED FP64 FLOP/s / peak (executed FLOPs per event from the committed ncu capture of this
           kernel, profiles/ncu_traffic.json); `frac_reference_arithmetic` = the reference's arithmetic
           (1329 + 2367 L_distinct + 226 L_cached FLOP/event, counted on the actual coszen array) / peak, which
           exceeds 1 because the kernel's formulation executes ~6.6x fewer FLOPs for the same result.
parity_check : the oracle chain (oracle/*.c) on a 1e5-event subsample of the bench's OWN events against the GPU
           path: bin indices bit-exact, binned weights relative.
cpu_baseline : the reference's own numba kernels (baseline/_ref, `parallel` target, all host cores) on a bounded
           sample; the C/OpenMP oracle port is timed beside it (`cpu_baseline_port`).
variants : measured in the same run at the current world size -- neutrino decay, C4 (standard NSI, 1.25e8 events/GPU), the FP32
           mode, the 40x40x2 stress binning, and C5 (100-point theta23 x dm31 scan with device-side mod_chi2).
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
    p.add_argument("--no-variants", action="store_true")
    p.add_argument("--no-parity", action="store_true")
    p.add_argument("--cpu-seconds", type=float, default=8.0)
    p.add_argument("--cpu-impl", default="auto", choices=["auto", "numba", "port"],
                   help="--impl reference: the reference's numba kernels (baseline/_ref) or the C oracle port")
    return p.parse_args()


def events_per_gpu(args):
    return int(args.events_per_gpu) // 12 * 12


def config_dict(args, world):
    """Identical in both arms (same workload, same numbers)."""
    n = events_per_gpu(args)
    return {
        "workload": "C3: synthetic per-event prob3 through PREM_12layer + weighted hist (sumw2), "
                    "dragon_datarelease 8x8x2, 12 flavour containers" + (", standard NSI" if args.nsi else ""),
        "events_per_gpu": n, "global_events": n * world, "n_bins": 128,
        "containers": 12, "earth_model": "PREM_12layer", "osc": "nufit v2.0 NH",
        "l2": "inputs (%.1f GB/GPU) larger than L2; no flush" % (n * 44 / 1e9),
        "parallelism": "events sharded over %d GPU(s); one all-reduce of [12,2,128] f64 per step" % world,
    }


# ------------------------------------------------------------------------------------------------
# CPU arms: the reference's own numba kernels (baseline/_ref) and the C oracle port
# ------------------------------------------------------------------------------------------------
def numba_reference(extra, timeout=600):
    """Run baseline/numba_chain.py (the UNMODIFIED reference numba path) in its own process; returns its JSON
    dict, or {"unavailable": why}."""
    from baseline import ref_pkg
    if not (ref_pkg.ref_built() or ref_pkg.reference_available()):
        return {"unavailable": "baseline/_ref is not built (run __graft_entry__.build() where /root/reference exists)"}
    try:
        import numba  # noqa: F401
    except ImportError:
        return {"unavailable": "numba is not installed on this machine"}
    cmd = [sys.executable, os.path.join(ROOT, "baseline", "numba_chain.py")] + [str(x) for x in extra]
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    except subprocess.TimeoutExpired:
        return {"unavailable": "numba_chain.py timed out after %d s" % timeout}
    for line in reversed(res.stdout.strip().splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    return {"unavailable": "numba_chain.py failed: %s" % res.stderr.strip()[-300:]}


def oracle_chain(oracle, L, mats, nubar, flav, ev, den=None, dis=None, n_threads=None, dims="dragon"):
    """Reference stage sequence on one container through the oracle: propagate_array -> fill_probs ->
    weights *= flux.prob -> bin index -> histogram (w, w^2) (prob3.py:581-622; hist.py:198-209).
    Returns (hist[2, 128], index)."""
    from pisa_b200.utils import synthetic as syn
    dm, mix, mat_pot = mats[:3]
    zc, zf = np.zeros((3, 3), dtype=np.complex128), np.zeros((3, 3))
    decay_flag = -1
    if len(mats) > 3:   # neutrino decay: (dm, mix, mat_pot, mat_decay)
        decay_flag, zc = 1, np.asarray(mats[3], dtype=np.complex128)
    if den is None:
        _, den, dis = L.calcLayers(ev["true_coszen"].astype(np.float64))
    prob = oracle.propagate_array(dm, mix, mat_pot, decay_flag, zc, zf, nubar, ev["true_energy"].astype(np.float64), den, dis,
                                  n_threads=n_threads or (os.cpu_count() or 1))
    pe, pmu = oracle.fill_probs(prob, 0, flav), oracle.fill_probs(prob, 1, flav)
    w = ev["weights"] * (ev["nu_flux"][:, 0] * pe + ev["nu_flux"][:, 1] * pmu)
    if dims == "stress":   # STRESS_DIMS: log-E 40 x lin-coszen 40 x pid 2 (the log of sample and domain, container.py:845-850)
        idx, _ = oracle.regular_index([np.log(ev["reco_energy"]), ev["reco_coszen"], ev["pid"]],
                                      [np.log(5.62341325), -1.0, -0.5], [np.log(56.23413252), 1.0, 1.5], [40, 40, 2])
        n_bins = 3200
    else:
        ie = oracle.digitize_irregular(ev["reco_energy"], syn.DRAGON_E_EDGES)
        i2, _ = oracle.regular_index([ev["reco_coszen"], ev["pid"]], [-1.0, -0.5], [1.0, 1.5], [8, 2])
        idx = np.where((ie >= 0) & (ie < 8) & (i2 >= 0), ie * 16 + i2, -1)
        n_bins = 128
    return np.stack([oracle.accumulate(idx, w, n_bins), oracle.accumulate(idx, w * w, n_bins)]), idx


class CpuChain:
    """The oracle port (C + OpenMP over events) of the reference's stage sequence.  Like the reference
    (prob3.setup_function, prob3.py:406-409) the Earth layers are computed once in setup, not per step."""

    def __init__(self, nsi=False):
        import oracle
        from pisa_b200.utils import synthetic as syn
        self.oracle, self.syn = oracle, syn
        self.threads = os.cpu_count() or 1
        prem = np.loadtxt(os.path.join(ROOT, "pisa_b200", "resources", "osc", "PREM_12layer.dat"))
        self.L = oracle.OracleLayers(prem, syn.EARTH["detector_depth"], syn.EARTH["prop_height"])
        self.L.setElecFrac(syn.EARTH["YeI"], syn.EARTH["YeO"], syn.EARTH["YeM"])
        self.mats = syn.osc_matrices(nsi=syn.STD_NSI if nsi else None)

    def make(self, n, seed=1):
        ev = self.syn.make_events_numpy(n, seed)
        per = n // 12
        self.blocks = []
        for c, (name, nubar, flav) in enumerate(self.syn.CONTAINERS):
            sl = slice(c * per, (c + 1) * per if c < 11 else n)
            b = {k: v[sl] for k, v in ev.items()}
            _, den, dis = self.L.calcLayers(b["true_coszen"])          # setup
            self.blocks.append((nubar, flav, b, den, dis))
        self.n = n

    def step(self):
        out = np.zeros((12, 2, 128))
        for c, (nubar, flav, ev, den, dis) in enumerate(self.blocks):
            out[c], _ = oracle_chain(self.oracle, self.L, self.mats, nubar, flav, ev, den, dis, self.threads)
        return out

    def sized_for(self, seconds):
        """events for ~`seconds` of CPU work per step (pilot run)."""
        self.make(24000)
        self.step()  # loads / warms the library
        t0 = time.perf_counter()
        self.step()
        rate = 24000 / (time.perf_counter() - t0)
        return int(min(max(rate * seconds, 48000), 2e7)) // 12 * 12, rate

    def timed(self, seconds, steps=1, warmup=0):
        n, _ = self.sized_for(seconds)
        self.make(n)
        for _ in range(warmup):
            self.step()
        t0 = time.perf_counter()
        for _ in range(steps):
            self.step()
        dt = time.perf_counter() - t0
        return {"value": n * steps / dt, "unit": UNIT, "cores": self.threads, "kind": "port",
                "sample": "%d synthetic events/step x %d step(s), oracle port (C + OpenMP, %d threads), layers in setup"
                          % (n, steps, self.threads)}, dt / steps


def numba_baseline_dict(r):
    return {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "reference",
            "sample": "%d synthetic events/step x %d step(s): the reference's own numba kernels (%s, target=%s, %s, "
                      "numba %s, %s threading), numpy.histogramdd for fast_histogram, layers in setup"
                      % (r["events"], r["steps"], r["origin"], r["target"], r["ftype"], r["numba"], r["threading_layer"]),
            "first_call_seconds_incl_jit": r["first_call_seconds"], "host_cpus": r["host_cpus"]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    per_step_s = max(1.0, min(15.0, 120.0 / max(1, args.steps + args.warmup)))
    cpu, note = None, "--cpu-impl port"
    if args.cpu_impl in ("auto", "numba"):
        extra = ["--target", "parallel", "--ftype", "fp64" if args.dtype == "f64" else "fp32", "--seconds-per-step",
                 per_step_s, "--steps", args.steps, "--warmup", max(1, args.warmup)] + (["--nsi"] if args.nsi else [])
        r = numba_reference(extra, timeout=900)
        if "unavailable" not in r:
            cpu = numba_baseline_dict(r)
            ms_step = 1e3 * r["seconds"] / r["steps"]
        elif args.cpu_impl == "numba":
            print(json.dumps({"impl": "reference", "unavailable": r["unavailable"]}), flush=True)
            return
        else:
            note = r["unavailable"]
    if cpu is None:
        cpu, s_step = CpuChain(nsi=args.nsi).timed(per_step_s, steps=args.steps, warmup=args.warmup)
        cpu["why_not_reference"] = note
        ms_step = 1e3 * s_step
    line = {
        "impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": config_dict(args, world), "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
STRESS_DIMS = [  # a `reco_binning`-like 40 x 40 x 2 = 3200-bin stress case on the same coordinate ranges
    dict(name="reco_energy", kind="log", n_bins=40, lo=5.62341325, hi=56.23413252),
    dict(name="reco_coszen", kind="lin", n_bins=40, lo=-1.0, hi=1.0),
    dict(name="pid", kind="lin", n_bins=2, lo=-0.5, hi=1.5),
]


class Harness:
    """Process-group plumbing, workload construction and device-side timing shared by the headline and the variants."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from pisa_b200 import distributed as D
        from pisa_b200 import ops
        from pisa_b200.stages.osc.layers import Layers
        from pisa_b200.utils import synthetic as syn
        self.torch, self.dist, self.ops, self.syn, self.args = torch, dist, ops, syn, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the native arm has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        # every rank holds its own slice of every container; the histogram exchange is on exactly when N > 1
        D.enable_event_sharding(self.world > 1)
        self.layers = Layers(os.path.join(ROOT, "pisa_b200", "resources", syn.EARTH["earth_model"]),
                             syn.EARTH["detector_depth"], syn.EARTH["prop_height"])
        self.layers.setElecFrac(syn.EARTH["YeI"], syn.EARTH["YeO"], syn.EARTH["YeM"])
        self.earth = self.layers.earth_struct()

    def consts(self, nsi=False, **osc):
        params = dict(self.syn.NUFIT20_NH, **osc)
        dm, mix, mat_pot = self.syn.osc_matrices(params, nsi=self.syn.STD_NSI if nsi else None)
        return self.ops.OscConsts.from_matrices(dm, mix, mat_pot)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t)

    def gather(self, x):
        if self.world == 1:
            return [float(x)]
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [float(v) for v in out]

    def build(self, n_rank, dtype, dims=None, n_bins=None, host=False, keep=0, seed_base=0):
        """Resident engine (and optionally a pinned-host twin) over 12 containers of n_rank / 12 synthetic events.
        ``keep`` > 0: also return the first `keep` events of every container (all columns + index) as numpy."""
        from pisa_b200.engine import ReweightEngine
        ops, syn, dev = self.ops, self.syn, self.dev
        dims = dims or syn.DRAGON_DIMS
        n_bins = n_bins or syn.DRAGON_NBINS
        binning, _keep = ops.make_binning(dims, dev)
        eng = ReweightEngine(self.earth, n_bins, dtype, dev)
        eng_host = ReweightEngine(self.earth, n_bins, dtype, dev) if host else None
        per = n_rank // 12
        sum_distinct = sum_cached = 0.0
        kept = []
        for c, (name, nubar, flav) in enumerate(syn.CONTAINERS):
            ev = syn.make_events_torch(per, seed=seed_base + 1000 * self.rank + c + 1, dtype=dtype, device=dev)
            # setup-time work, like hist.setup_function / Container.translate: the static bin index
            index = ops.hist_index(binning, [ev["reco_energy"], ev["reco_coszen"], ev["pid"]])
            _, distinct, cached = syn.layer_counts(self.layers.coszen_limit, ev["true_coszen"])
            sum_distinct += float(distinct.sum())
            sum_cached += float(cached.sum())
            arrays = dict(true_energy=ev["true_energy"], true_coszen=ev["true_coszen"], nu_flux=ev["nu_flux"],
                          weights=ev["weights"], index=index)
            if keep:
                sub = {k: v[:keep].cpu().numpy() for k, v in ev.items()}
                sub["index"] = index[:keep].cpu().numpy()
                kept.append((name, nubar, flav, sub))
            eng.add_container(name, nubar, flav, **arrays)
            if host:
                eng_host.add_container(name, nubar, flav, **{k: v.cpu() for k, v in arrays.items()})
            del ev
        self.torch.cuda.synchronize()
        return eng, eng_host, syn.flops_per_event(sum_distinct / max(n_rank, 1), sum_cached / max(n_rank, 1)), kept

    def time_steps(self, fn, steps, warmup, nvtx=None):
        """W warm-up calls, then K calls bracketed by barrier + synchronize on both sides, timed with CUDA events on
        the launching stream; returns ms per step (max over ranks) and the last result."""
        torch = self.torch
        for _ in range(warmup):
            out = fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        if nvtx:
            torch.cuda.nvtx.range_push(nvtx)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        self.barrier()
        if nvtx:
            torch.cuda.nvtx.range_pop()
        return self.max_over_ranks(e0.elapsed_time(e1)) / steps, out


def run_variants(h, args):
    """BASELINE configs C4 / C5, the FP32 mode and the stress binning, at the current world size."""
    torch, ops, syn = h.torch, h.ops, h.syn
    steps, warmup = max(2, min(args.steps, 3)), 3
    out = {}
    consts = h.consts()

    # ---- C4: standard NSI + hist, 1e9 events over 8 GPUs = 1.25e8 events per GPU (weak scaling) -------------
    n = 125_000_000 // 12 * 12
    keep = 0 if (args.no_parity or h.rank != 0) else 2000
    eng, _, _, kept = h.build(n, np.float64, seed_base=50_000, keep=keep)
    c_nsi = h.consts(nsi=True)
    ms, res = h.time_steps(lambda: eng.evaluate(c_nsi), steps, warmup)
    par = parity_check(h, kept, syn.osc_matrices(nsi=syn.STD_NSI), lambda e: e.evaluate(c_nsi, allreduce=False)) if kept else None
    out["C4_nsi"] = {"value": n * h.world / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "events_per_gpu": n,
                     "parity_check": par,
                     "global_events": n * h.world, "dtype": "f64", "n_bins": 128, "steps": steps, "warmup": warmup,
                     "what": "BASELINE configs[3]: standard-NSI matter potential (eps_emu 0.07/340deg, eps_etau "
                             "0.06/35deg, eps_mutau 0.003/175deg) + weighted hist; general-Hamiltonian kernel path"}
    del eng, res
    torch.cuda.empty_cache()

    # ---- FP32 mode (BASELINE configs[2] "FP64 and FP32"), same events rounded to float32 ----------------------
    n = events_per_gpu(args)
    eng64, _, _, _ = h.build(n, np.float64)
    ref64 = eng64.evaluate(consts).clone()
    del eng64
    torch.cuda.empty_cache()
    eng32, _, _, _ = h.build(n, np.float32)
    ops.set_f32_math("mixed")
    ms, res = h.time_steps(lambda: eng32.evaluate(consts), steps, warmup)
    nz = ref64[:, 0] > 0
    rel = float(((res[:, 0] - ref64[:, 0]).abs()[nz] / ref64[:, 0][nz]).max())
    ops.set_f32_math("fp64")
    ms_s, _ = h.time_steps(lambda: eng32.evaluate(consts), steps, warmup)
    ops.set_f32_math(args.f32_math)
    out["f32"] = {"value": n * h.world / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "events_per_gpu": n,
                  "dtype": "f32", "math": "mixed: FP64 eigenvalues / phase arguments / geometry, float32 matrices, "
                                          "products and state (csrc/prob3_mp.cuh)",
                  "max_rel_diff_binned_weights_vs_f64_run": rel,
                  "max_rel_diff_note": "includes events that the float32 rounding of the reco coordinates moves across a "
                                       "bin edge (one event in ~1e4 per bin); the arithmetic alone: <= 1e-5 per "
                                       "probability, tests/test_gpu_prob3.py",
                  "value_f32_storage_fp64_math": n * h.world / (ms_s * 1e-3), "steps": steps, "warmup": warmup}
    del eng32, res
    torch.cuda.empty_cache()

    # ---- stress binning 40 x 40 x 2 = 3200 bins ----------------------------------------------------
    eng, _, _, kept = h.build(n, np.float64, dims=STRESS_DIMS, n_bins=3200, keep=keep)
    ms, res = h.time_steps(lambda: eng.evaluate(consts), steps, warmup)
    again = eng.evaluate(consts).clone()
    again2 = eng.evaluate(consts)
    par = parity_check(h, kept, syn.osc_matrices(), lambda e: e.evaluate(consts, allreduce=False), dims="stress") if kept else None
    out["bins_3200"] = {"value": n * h.world / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "events_per_gpu": n,
                        "parity_check": par,
                        "n_bins": 3200, "binning": "log-E 40 x lin-coszen 40 x pid 2", "dtype": "f64",
                        "bit_reproducible_run_to_run": bool(torch.equal(again, again2)), "steps": steps,
                        "warmup": warmup}
    del eng, res, again, again2
    torch.cuda.empty_cache()

    # ---- neutrino decay (prob3 neutrino_decay=True: the reference's numpy.linalg.eigvals branch) -------------------
    eng, _, _, kept = h.build(n, np.float64, seed_base=90_000, keep=keep)
    dm_, mix_, mp_ = syn.osc_matrices()
    mat_decay = np.zeros((3, 3), dtype=np.complex128)
    mat_decay[2, 2] = -1.0e-4j   # decay_params.py:47-55 with alpha3 = 1e-4 eV^2
    c_dec = ops.OscConsts.from_matrices(dm_, mix_, mp_, 1, mat_decay)
    ms, res = h.time_steps(lambda: eng.evaluate(c_dec), steps, warmup)
    par = parity_check(h, kept, (dm_, mix_, mp_, mat_decay), lambda e: e.evaluate(c_dec, allreduce=False)) if kept else None
    out["decay"] = {"value": n * h.world / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "events_per_gpu": n,
                    "parity_check": par, "dtype": "f64", "n_bins": 128, "steps": steps, "warmup": warmup,
                    "what": "osc.prob3 with neutrino_decay=True (decay_flag = 1, alpha3 = 1e-4 eV^2): non-Hermitian layer "
                            "Hamiltonian, general-matrix kernels (csrc/prob3_decay.cuh); the reference solves it with "
                            "numpy.linalg.eigvals per layer on the CPU"}
    del eng, res
    torch.cuda.empty_cache()

    # ---- C5: theta23 x dm31 scan, mod_chi2 on device, every point a full template --------------------------
    from pisa_b200 import scan
    fixed = dict(theta12=np.deg2rad(33.48), theta13=np.deg2rad(8.5), deltacp=0.0, dm21=7.5e-5)
    t23 = np.deg2rad(np.linspace(38.0, 52.0, 10))
    dm31 = np.linspace(2.2e-3, 2.7e-3, 10)
    points = [(a, b) for a in t23 for b in dm31]
    out["C5_scan"] = {"points": len(points), "what": "BASELINE configs[4]: 10 x 10 theta23 x dm31 grid, each point = "
                      "prob3 + reweight + hist over ALL events of the sample + the histogram exchange + mod_chi2 on "
                      "the device; sample sharded over the ranks (strong scaling), all hypotheses of the grid in one "
                      "launch per rank", "sizes": []}
    for n_total in (120_000, 12_000_000):
        n_rank = n_total // h.world // 12 * 12
        eng, _, _, _ = h.build(n_rank, np.float64, seed_base=70_000)
        observed = scan.asimov(eng, h.consts(theta23=45.0))

        def one_scan():
            return scan.scan_chi2(eng, observed, points, fixed, batch=len(points))
        ms, chi2 = h.time_steps(one_scan, 3, 2)

        def seq_scan():
            return scan.scan_chi2(eng, observed, points[:20], fixed, batch=1)
        ms_seq, _ = h.time_steps(seq_scan, 2, 1)
        chi2 = chi2.cpu().numpy()
        out["C5_scan"]["sizes"].append({
            "events_per_template": n_rank * h.world, "templates_per_s": len(points) / (ms * 1e-3),
            "events_per_s": n_rank * h.world * len(points) / (ms * 1e-3), "ms_per_scan": ms,
            "templates_per_s_one_launch_per_template": 20 / (ms_seq * 1e-3),
            "chi2_min": float(chi2.min()), "argmin": int(chi2.argmin())})
        del eng
        torch.cuda.empty_cache()
    return out


def parity_check(h, kept, consts_mats, eng_out_fn, dims="dragon"):
    """Oracle chain on the kept subsample of the bench's own events vs the GPU path on the same events."""
    import oracle
    from pisa_b200.engine import ReweightEngine
    torch, syn, dev = h.torch, h.syn, h.dev
    prem = np.loadtxt(os.path.join(ROOT, "pisa_b200", "resources", "osc", "PREM_12layer.dat"))
    OL = oracle.OracleLayers(prem, syn.EARTH["detector_depth"], syn.EARTH["prop_height"])
    OL.setElecFrac(syn.EARTH["YeI"], syn.EARTH["YeO"], syn.EARTH["YeM"])
    n_bins = 3200 if dims == "stress" else syn.DRAGON_NBINS
    eng = ReweightEngine(h.earth, n_bins, kept[0][3]["weights"].dtype, dev)
    ref = np.zeros((len(kept), 2, n_bins))
    mism = 0
    for c, (name, nubar, flav, ev) in enumerate(kept):
        ref[c], idx = oracle_chain(oracle, OL, consts_mats, nubar, flav, ev, dims=dims)
        mism += int((idx.astype(np.int32) != ev["index"]).sum())
        t = {k: torch.tensor(ev[k], device=dev) for k in ("true_energy", "true_coszen", "nu_flux", "weights", "index")}
        eng.add_container(name, nubar, flav, **t)
    got = eng_out_fn(eng).cpu().numpy()
    nz = ref != 0
    rel = np.abs(got - ref)[nz] / np.abs(ref)[nz]
    n = sum(len(k[3]["weights"]) for k in kept)
    return {"events": n, "oracle": "oracle/*.c (parity pinned to the reference's golden vectors)",
            "index_mismatches": mism, "max_rel_err_sum_w": float(rel.reshape(-1).max()) if rel.size else 0.0,
            "tolerance": 1e-10 if got.dtype == np.float64 and kept[0][3]["weights"].dtype == np.float64 else 1e-4,
            "ok": bool(mism == 0 and (rel.size == 0 or rel.max() <= (1e-10 if kept[0][3]["weights"].dtype == np.float64 else 1e-4)))}


def run_native(args):
    h = Harness(args)
    torch, ops, syn = h.torch, h.ops, h.syn
    rank, world, dev = h.rank, h.world, h.dev
    dtype = np.float64 if args.dtype == "f64" else np.float32
    ops.set_f32_math(args.f32_math)
    n_gpu = events_per_gpu(args)
    consts = h.consts(nsi=args.nsi)
    eng, eng_host, flops_event, kept = h.build(n_gpu, dtype, host=not args.no_e2e,
                                               keep=0 if args.no_parity else 8334)

    # ---- resident-input steps ------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        eng.evaluate(consts)
    sampler = ClockSampler(h.local)
    h.barrier()
    sampler.start()
    ops.launch_count(reset=True)
    kernel_events = []
    ms_step, out = h.time_steps(lambda: eng.evaluate(consts, events=kernel_events), args.steps, 0, nvtx="timed")
    launches = ops.launch_count()
    clocks = sampler.stop()
    value = n_gpu * world / (ms_step * 1e-3)
    # per-launch CUDA-event times bracket the fused kernel (ONE launch for all 12 containers) + its tiny
    # partial-reduction kernel
    launch_ms = float(np.mean([a.elapsed_time(b) for a, b in kernel_events]))
    hist_total = float(out[:, 0].sum())

    # ---- FP64 roofline ---------------------------------------------------------------------------
    peak_flops = max(ops.fp64_peak_probe(20000)[0] for _ in range(3))   # best of 3: the first probe can catch a clock ramp
    kname = "reweight_hist_kernel<%s>" % ("double" if args.dtype == "f64" else "float")
    traffic = pipe_active = executed_flop = tr_src = None
    tr = {}
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tr = json.load(f)[kname]
        traffic = tr["dram_bytes_per_event"] * n_gpu
        pipe_active = tr.get("fp64_pipe_active_pct")
        executed_flop = tr.get("executed_fp64_flop_per_event")
        tr_src = tr.get("source")
    except (OSError, KeyError, ValueError):
        pass
    ref_arith = flops_event * n_gpu / (launch_ms * 1e-3)
    executed = None if executed_flop is None else executed_flop * n_gpu / (launch_ms * 1e-3)
    bytes_event = 44 if args.dtype == "f64" else 24
    roofline = {
        "bound": "fp64", "kernel": kname,
        "achieved": None if executed is None else executed / 1e12, "peak": peak_flops / 1e12, "unit": "TFLOP/s",
        "frac": None if executed is None else executed / peak_flops,
        "frac_is": "EXECUTED FP64 FLOP/s (2 DFMA + DMUL + DADD per event from the ncu capture below x events / "
                   "CUDA-event launch time) / measured DFMA peak",
        "traffic": traffic,
        "peak_source": "DFMA micro-benchmark in this process (MEASURED_PEAKS.json has no FP64 entry); "
                       "nominal %.1f TFLOP/s" % NOMINAL_FP64_TFLOPS,
        "executed_flops_per_event_ncu": executed_flop, "fp64_pipe_active_pct_ncu": pipe_active,
        "ncu_source": tr_src, "events_per_launch": n_gpu, "launch_ms": launch_ms,
        "algorithmic_flops_per_event": flops_event, "achieved_reference_arithmetic": ref_arith / 1e12,
        "frac_reference_arithmetic": ref_arith / peak_flops,
        "frac_reference_arithmetic_is": "SURVEY 8d contract: the REFERENCE's FLOPs per event x events / launch time / "
                                        "peak; > 1 because this kernel's formulation executes ~6.6x fewer FLOPs",
        "algorithmic_bytes_per_event": bytes_event, "hbm_gbs": bytes_event * n_gpu / (launch_ms * 1e-3) / 1e9,
    }

    if args.dtype == "f32" and args.f32_math == "mixed" and tr.get("warp_instructions_per_warp_event"):
        # the FP32 mode is bound by the ISSUE port (one warp instruction per cycle per SM sub-partition;
        # profiles/r02_pipe_mix.txt), not by the FP64 pipe: report that fraction as `frac`
        sms = ops.device_info()["sm_count"]
        clk = (clocks.get("sm_mhz") or 1965.0) * 1e6
        issued = tr["warp_instructions_per_warp_event"] * (n_gpu / 32.0) / (launch_ms * 1e-3) / (sms * 4 * clk)
        roofline.update({
            "bound": "issue", "achieved": issued, "peak": 1.0, "unit": "warp-instructions/cycle/SM sub-partition",
            "frac": issued, "frac_is": "executed warp instructions per event (ncu capture below) x events / launch time / "
                                       "(SMs x 4 sub-partitions x SM clock): the issue-port utilisation",
            "issue_active_pct_ncu": tr.get("issue_active_pct"),
            "executed_fp32_flops_per_event_ncu": tr.get("executed_fp32_flop_per_event"),
            "executed_fp64_tflops": None if executed is None else executed / 1e12,
        })

    # ---- end to end (host buffers) -------------------------------------------------------------
    e2e = e2e_changed = None
    if not args.no_e2e:
        def e2e_run(changed):
            for _ in range(2):
                eng_host.evaluate_host(consts, changed=changed)
            h.barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                host_out = eng_host.evaluate_host(consts, changed=changed)
            torch.cuda.synchronize()
            t_mine = time.perf_counter() - t0
            t_all = h.max_over_ranks(t_mine)
            h.barrier()
            per_rank = h.gather(eng_host.last_h2d_bytes * args.steps / t_mine / 1e9)
            return host_out, t_all, per_rank
        host_out, t_e2e, per_rank = e2e_run(None)
        h2d_full, d2h = int(eng_host.last_h2d_bytes), int(eng_host.last_d2h_bytes)
        if world == 1 and abs(float(host_out[:, 0].sum()) / hist_total - 1) > 1e-9:
            raise SystemExit("bench.py: e2e and resident histograms disagree")
        # raw pinned-host -> device copy bandwidth of this box, as the yardstick for the e2e number
        probe = torch.empty(1 << 28, dtype=torch.uint8).pin_memory()
        probe_d = torch.empty(1 << 28, dtype=torch.uint8, device=dev)
        probe_d.copy_(probe, non_blocking=True)
        h.barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(4):
            probe_d.copy_(probe, non_blocking=True)
        p1.record()
        torch.cuda.synchronize()
        h2d_peak = h.gather(4 * probe.numel() / (p0.elapsed_time(p1) * 1e-3) / 1e9)
        del probe, probe_d
        e2e = {"value": n_gpu * world * args.steps / t_e2e, "unit": UNIT,
               "h2d_bytes_per_step": h2d_full, "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * t_e2e / args.steps,
               "mode": "every event array travels on every step (energy, coszen, nu_flux, weights, bin index)",
               "h2d_gbs": h2d_full * args.steps / t_e2e / 1e9, "h2d_gbs_per_rank": per_rank,
               "h2d_peak_gbs_measured_per_rank_concurrent": h2d_peak, "limit": "host -> device copies (PCIe / host memory)",
               "api": "pisa_b200.engine.ReweightEngine.evaluate_host (pinned host arrays, double-buffered H2D)"}
        host_out2, t2, per_rank2 = e2e_run(("weights", "nu_flux"))
        e2e_changed = {"value": n_gpu * world * args.steps / t2, "unit": UNIT,
                       "h2d_bytes_per_step": int(eng_host.last_h2d_bytes), "d2h_bytes_per_step": d2h,
                       "ms_per_step": 1e3 * t2 / args.steps, "h2d_gbs_per_rank": per_rank2,
                       "mode": "only the arrays a fit changes between templates travel (weights, nu_flux: 24 B/event in "
                               "FP64); energy, coszen and the static bin index were uploaded once",
                       "api": "ReweightEngine.evaluate_host(changed=('weights', 'nu_flux'))"}
        if world == 1 and abs(float(host_out2[:, 0].sum()) / hist_total - 1) > 1e-9:
            raise SystemExit("bench.py: changed-only e2e and resident histograms disagree")
        del eng_host
    parity = None
    if kept and rank == 0:
        mats = syn.osc_matrices(nsi=syn.STD_NSI if args.nsi else None)
        parity = parity_check(h, kept, mats, lambda e: e.evaluate(consts, allreduce=False))
    del eng, out
    torch.cuda.empty_cache()

    # ---- variants (all ranks take part) ----------------------------------------------------------
    variants = None
    if not args.no_variants:
        variants = run_variants(h, args)

    # ---- CPU baselines on rank 0, N = 1 only ---------------------------------------------------
    cpu = cpu_port = cpu_numba = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_port, _ = CpuChain(nsi=args.nsi).timed(args.cpu_seconds)
        nsi = ["--nsi"] if args.nsi else []
        runs = {"fp64_parallel": ["--target", "parallel", "--ftype", "fp64", "--seconds-per-step", args.cpu_seconds],
                "fp32_parallel": ["--target", "parallel", "--ftype", "fp32", "--seconds-per-step", args.cpu_seconds / 2],
                "fp64_cpu_1core": ["--target", "cpu", "--ftype", "fp64", "--seconds-per-step", args.cpu_seconds / 2]}
        cpu_numba = {}
        for k, extra in runs.items():
            r = numba_reference(extra + nsi)
            cpu_numba[k] = r if "unavailable" in r else numba_baseline_dict(r)
        main = cpu_numba["fp64_parallel" if args.dtype == "f64" else "fp32_parallel"]
        cpu = main if "unavailable" not in main else cpu_port

    if world > 1:
        h.dist.destroy_process_group()
    if rank == 0:
        return {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "config": config_dict(args, world),
            "clocks": clocks, "e2e": e2e, "e2e_changed_only": e2e_changed, "gpu_launches": int(launches),
            "roofline": roofline, "parity_check": parity, "cpu_baseline": cpu, "cpu_baseline_port": cpu_port,
            "cpu_baseline_numba": cpu_numba, "variants": variants,
        }
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
