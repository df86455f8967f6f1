"""layers_kernel (Layers.calcLayers on the device): GB/s of the [n, max_layers] outputs at a size far above L2."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pisa_b200 import ops
from pisa_b200.stages.osc.layers import Layers
dev = torch.device("cuda:0")
for model, n in (("PREM_12layer.dat", 20_000_000), ("PREM_59layer.dat", 4_000_000), ("PREM_4layer.dat", 40_000_000)):
    L = Layers(os.path.join(ROOT, "pisa_b200/resources/osc", model), 2.0, 20.0); L.setElecFrac(0.4656, 0.4656, 0.4957)
    earth = L.earth_struct()
    for dt in (torch.float64, torch.float32):
        g = torch.Generator(device=dev); g.manual_seed(1)
        cz = (torch.rand(n, generator=g, device=dev, dtype=torch.float64) * 2 - 1).to(dt)
        ops.layers_calc(earth, cz); torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); out = ops.layers_calc(earth, cz); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        es = cz.element_size()
        bpe = es + 2 * es * earth.max_layers + 4
        print("%-16s %s %9d events: %8.3f ms  %7.1f GB/s (%d B/event)  %.2e events/s" % (model, str(dt)[6:], n, best, bpe * n / best / 1e6, bpe, n / best * 1e3), flush=True)
        del out
