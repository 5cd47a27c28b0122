"""Dev probe: train-step time of the BASELINE model configs with the Jacobian rows built by ONE batched (vmapped)
backward pass vs k sequential passes, eager and under a CUDA graph.   python tools/step_probe.py [configs...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import torch  # noqa: E402

import movae_b200  # noqa: E402,F401
from movae_b200 import autojac  # noqa: E402
import vqvae_harness as H  # noqa: E402

dev = torch.device("cuda", 0)
which = sys.argv[1:] or ["vqvae", "vae", "gg", "vq2"]
arms = ("movae_eager", "movae_graph")
for batched in (True, False):
    autojac.BATCHED_JACOBIAN = batched
    for name in which:
        if name == "vqvae":
            r = H.time_config(dev, lambda mq: H.VQVAEShell(mq(512, 64)), 128, 32, "aligned_mtl", 20, 5, arms)
        elif name == "gg":
            r = H.time_config(dev, lambda mq: H.GGVQVAEShell(mq(512, 64)), 256, 64, "mgda_lgn", 10, 3, arms)
        elif name == "vq2":
            r = H.time_config(dev, lambda mq: H.VQVAE2Shell(mq), 64, 256, "upgrad", 5, 2, arms)
        else:
            class _VAE(H.VAEShell):
                def forward(self, x):
                    f, l = super().forward(x)
                    return f, l, None
            r = H.time_config(dev, lambda mq: _VAE(), 128, 32, "upgrad", 20, 5, arms)
        print(json.dumps({"config": name, "batched_jacobian": batched, **r}), flush=True)
