#!/usr/bin/env python
"""One eval forward (+ decode) without warm-up, for ncu captures:
  ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 37 -c 1 -o gpurun_out/prof python tools/prof_forward.py 64
conv_igemm launch #38 of a forward is the fused 8-head conv1 (128 -> 1024, 3x3, 128x128)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import abcnet_b200  # noqa: E402
import synthdata as synth  # noqa: E402  (deterministic synthetic weights / images / targets; the oracle is not used here)
import synthdata as unet_ref  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
model = abcnet_b200.UNet(1, list(unet_ref.V2_HEADS)).cuda().eval()
model.load_state_dict(unet_ref.make_state_dict(0))
x = torch.from_numpy(synth.binary_images(0, 8, 512, 512, 0.05)).repeat((B + 7) // 8, 1, 1, 1)[:B].contiguous().cuda()
outs = model.infer(x, layout="p8f")
dec = abcnet_b200.PeakDecoder(B, 1024, 4096)
dec.launch(outs)
torch.cuda.synchronize()
print("done", abcnet_b200.launch_count())
