"""One eager force step between cudaProfilerStart/Stop, for `ncu --profile-from-start off` (see tools/ncu_step_summary.py).
    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r02_step python tools/prof_step.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

g.load_package()
from deepmd_kit_b200.model import DeepPotB200, SeAConfig, SeAModel  # noqa: E402

ncopy = int(os.environ.get("NCOPY", "8"))
dtype = torch.float64 if os.environ.get("DTYPE", "f64") == "f64" else torch.float32
dev = torch.device("cuda:0")
coord, atype, box = g.water_box(ncopy, 0.01)
if os.environ.get("MODEL", "se_a") == "se_atten":  # config 5 model (ncu summary of the gated kernels)
    from deepmd_kit_b200.atten import SeAttenConfig, SeAttenModel  # noqa: E402

    model = SeAttenModel(SeAttenConfig(), dtype, dev)
elif os.environ.get("MODEL") == "dpa1_attn":  # DPA-1 with two attention layers
    from deepmd_kit_b200.atten import SeAttenConfig, SeAttenModel  # noqa: E402

    model = SeAttenModel(SeAttenConfig(attn_layer=2), dtype, dev)
else:
    model = SeAModel(SeAConfig(), dtype, dev)
dp = DeepPotB200(model, skin=2.0, nlist_every=10, use_graph=False)
c = torch.as_tensor(coord.astype(np.float64 if dtype == torch.float64 else np.float32)).to(dev)
t = torch.as_tensor(atype).to(dev)
for _ in range(3):
    dp.eval_device(c, t, box)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
out = dp.eval_device(c, t, box)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("natoms", len(atype), "energy", float(out[0]))
