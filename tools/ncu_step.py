"""One eager training iteration (cfg3 by default) bracketed by cudaProfilerStart/Stop, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum[,dram__bytes_*.sum] --clock-control none ...
(the launch list committed under profiles/: per-kernel shares of the step, DRAM traffic per launch)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
from cutmix_semisup_seg_b200 import synthetic  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else 'v3plus'
cfg = bench.CFG[arch]
dev = torch.device('cuda:0')
torch.cuda.set_device(dev)
trainer, mg = bench.build_trainer(cfg, dev, False, use_graph=False)
n, h, w = cfg['batch'], cfg['h'], cfg['w']
sup = synthetic.make_sup_batch(n, h, w, cfg['classes'], 100)
uns = synthetic.make_unsup_batch(n, h, w, 200, mg)
sup = tuple(t.to(dev) for t in sup)
uns = {k: v.to(dev) for k, v in uns.items()}
for _ in range(2):
    trainer.step(sup, [uns])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
trainer.step(sup, [uns])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
