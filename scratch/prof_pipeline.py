import cProfile, pstats, os, sys, io
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pisa_b200.core.pipeline import Pipeline
from pisa_b200.utils.units import ureg
for cfg in sys.argv[1:]:
    pipe = Pipeline(cfg); pipe.run(); pipe.get_outputs(); torch.cuda.synchronize()
    pr = cProfile.Profile(); pr.enable()
    th = 42.3
    for i in range(5):
        th += 0.3; pipe.params.theta23 = th * ureg.deg
        pipe.run(); pipe.get_outputs(); torch.cuda.synchronize()
    pr.disable()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
    print("=====", cfg); print("\n".join(l[:150] for l in s.getvalue().splitlines()[:60]))
