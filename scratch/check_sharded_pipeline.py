"""2-GPU check of the Stage API with events sharded over ranks (run under torchrun on `gpurun --gpus 2`):
the MapSet of a sharded Pipeline / FusedPipeline on every rank equals the single-process result."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pisa_b200 import distributed as D
rank = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(rank)
D.init_from_env(device=torch.device("cuda", rank))
from pisa_b200.core.pipeline import Pipeline
from pisa_b200.fused import FusedPipeline
from pisa_b200.utils.units import ureg
cfgs = ["settings/pipeline/b200_events.cfg", "settings/pipeline/b200_icecube3y_full.cfg"]
ok = True
for cfg in cfgs:
    D.enable_event_sharding(False)
    ref = Pipeline(cfg)
    ref.params.theta23 = 46.0 * ureg.deg
    ref_maps = ref.get_outputs()
    D.enable_event_sharding(True)
    assert D.event_sharding()
    sharded = Pipeline(cfg)
    def count(pipe):
        total = 0
        for c in pipe.data.containers:
            c.representation = "events"
            total += int(c["true_energy"].shape[0])
        return total
    n_local = count(sharded)
    sharded.params.theta23 = 46.0 * ureg.deg
    maps = sharded.get_outputs()
    fused = FusedPipeline(Pipeline(cfg))
    fused.pipeline.params.theta23 = 46.0 * ureg.deg
    fmaps = fused.get_outputs()
    worst = 0.0
    for m in ref_maps:
        for other in (maps, fmaps):
            worst = max(worst, float(np.abs(other[m.name].hist / m.hist - 1).max()),
                        float(np.abs(other[m.name].std_devs / m.std_devs - 1).max()))
    n_all = count(ref)
    print("rank %d %s: local events %d of %d, max rel deviation of maps and errors %.2e" % (rank, cfg.split("/")[-1], n_local, n_all, worst), flush=True)
    ok = ok and worst < 1e-10 and n_local < n_all
torch.distributed.barrier()
print("rank %d %s" % (rank, "SHARDED-OK" if ok else "SHARDED-FAIL"), flush=True)
torch.distributed.destroy_process_group()
