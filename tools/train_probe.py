"""Train-step timing of the BASELINE model configs (dev tool): python tools/train_probe.py [vqvae|vae|ggvqvae|vqvae2 ...]"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import vqvae_harness as H
dev = torch.device("cuda")
fns = {"vqvae": H.time_train_steps, "vae": H.time_vae_train_steps, "ggvqvae": H.time_ggvqvae_train_steps, "vqvae2": H.time_vqvae2_train_steps}
for name in (sys.argv[1:] or ["vqvae"]):
    r = fns[name](dev)
    print(name, json.dumps({k: (v["steps_per_s"] if isinstance(v, dict) and "steps_per_s" in v else v) for k, v in r.items()}))
