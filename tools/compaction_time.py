"""bench.py's valid-row compaction sweep on its own (cfg2 shape, right-padded captions): module path and C-ABI path."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import preference_guided_image_captioning_alignment_b200 as pg

dev = torch.device("cuda")
W = (torch.randn(bench.CFG["vocab"], bench.CFG["d"], generator=torch.Generator().manual_seed(1)) * 0.02).to(dev)
for rep in range(1):
    out = bench.compaction_extra(torch, pg, dev, W)
    print(json.dumps([[r["scored_rows"], round(r["module_ms_per_step"], 3), round(r["c_abi_ms_per_step"], 3)]
                      for r in out["sweep"]]), "speedup", round(out["speedup"], 3), flush=True)
