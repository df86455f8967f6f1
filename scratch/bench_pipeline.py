"""Pipeline-level timings of the drop-in boundary: Pipeline(cfg) setup, and run() per template when one
oscillation parameter changes (so osc.prob3 recomputes and everything downstream re-applies), like an
iteration of a fit.  Reference figures for the same stage chain (unstated CPU):
prob3 compute 0.887 s, prob3 apply 0.018 s, hist apply 0.034 s, run() 0.525 s mean / 0.040 s all caches hit."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pisa_b200.core.pipeline import Pipeline
from pisa_b200.utils.units import ureg

def bench(cfg, n_events=None, reps=10):
    t0 = time.perf_counter()
    pipe = Pipeline(cfg)
    if n_events is not None:
        for st in pipe.stages:
            if "n_events" in st.params.names:
                st.params.n_events = n_events * ureg.dimensionless
    pipe.setup() if hasattr(pipe, "setup") else None
    pipe.run(); torch.cuda.synchronize()
    t_setup = time.perf_counter() - t0
    th = 42.3
    ts = []
    for i in range(reps):
        th += 0.37
        pipe.params.theta23 = th * ureg.deg
        torch.cuda.synchronize(); t1 = time.perf_counter()
        pipe.run()
        out = pipe.get_outputs() if hasattr(pipe, "get_outputs") else None
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t1)
    n = 0
    for c in pipe.data.containers:
        try:
            c.representation = "events"
            n += int(c["true_energy"].shape[0])
        except Exception:
            pass
    print("%-48s events %9d  first run incl. setup %7.1f ms   run()+get_outputs per template: median %7.2f ms  min %7.2f ms"
          % (cfg.split("/")[-1] + ("" if n_events is None else " n=%g" % n_events), n, 1e3 * t_setup, 1e3 * np.median(ts), 1e3 * min(ts)), flush=True)

def bench_fused(cfg, n_events, reps=20):
    from pisa_b200.fused import FusedPipeline
    pipe = Pipeline(cfg)
    for st in pipe.stages:
        if "n_events" in st.params.names:
            st.params.n_events = n_events * ureg.dimensionless
    pipe.setup()
    fp = FusedPipeline(pipe)
    fp.get_outputs(); torch.cuda.synchronize()
    th, ts = 42.3, []
    for i in range(reps):
        th += 0.37
        pipe.params.theta23 = th * ureg.deg
        torch.cuda.synchronize(); t1 = time.perf_counter()
        fp.get_outputs()
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t1)
    print("%-48s FusedPipeline.get_outputs per template: median %7.2f ms  min %7.2f ms"
          % (cfg.split("/")[-1] + " n=%g" % n_events, 1e3 * np.median(ts), 1e3 * min(ts)), flush=True)

bench("settings/pipeline/b200_oscillogram.cfg")
for n in (20000, 1000000):
    bench("settings/pipeline/b200_events.cfg", n)
    bench("settings/pipeline/b200_icecube3y_like.cfg", n)
    bench("settings/pipeline/b200_icecube3y_events.cfg", n)
for n in (20000, 1000000):
    bench_fused("settings/pipeline/b200_events.cfg", n)
    bench_fused("settings/pipeline/b200_icecube3y_events.cfg", n)
