"""Bench plumbing (NOT part of the product path): a VQ-VAE model shell with the layer shapes of the
reference's `VQVAE` at the BASELINE config (models/vq_vae.py:148-318 with hidden_dims [128, 256],
2 residual layers, K = 512, D = 64: 2,448,064 shared encoder parameters, 1,988,995 decoder parameters),
written here because /root/reference does not exist on the GPU box.  Convolutions are plain torch.nn
(cuDNN); the quantizer and the multi-objective backward are the product's.

train_step() mirrors the reference's train_epoch body (main.py:157-214): zero_grad -> forward -> loss dict
[reconstruction, embedding, commitment] (vq_vae.py:185, :380-390, lambda 1 / 1 / 0.25) ->
mtl_backward(losses, features=[encoding], aggregator) -> Adam step.
"""
from __future__ import annotations

import torch
from torch import nn
import torch.nn.functional as F


class _Residual(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.block = nn.Sequential(nn.Conv2d(c, c, 3, padding=1, bias=False), nn.ReLU(True), nn.Conv2d(c, c, 1, bias=False))

    def forward(self, x):
        return x + self.block(x)


class VQVAEShell(nn.Module):
    def __init__(self, quantizer: nn.Module, in_channels=3, hidden=(128, 256), n_res=2, D=64):
        super().__init__()
        enc, c = [], in_channels
        for h in hidden:
            enc += [nn.Conv2d(c, h, 4, 2, 1), nn.LeakyReLU()]
            c = h
        enc += [nn.Conv2d(c, c, 3, 1, 1), nn.LeakyReLU()]
        enc += [_Residual(c) for _ in range(n_res)]
        enc += [nn.LeakyReLU(), nn.Conv2d(c, D, 1), nn.LeakyReLU()]
        self.encoder = nn.Sequential(*enc)
        self.vq_layer = quantizer
        dec = [nn.Conv2d(D, c, 3, 1, 1), nn.LeakyReLU()]
        dec += [_Residual(c) for _ in range(n_res)]
        dec += [nn.LeakyReLU()]
        rev = list(hidden)[::-1]
        for a, b in zip(rev[:-1], rev[1:]):
            dec += [nn.ConvTranspose2d(a, b, 4, 2, 1), nn.LeakyReLU()]
        dec += [nn.ConvTranspose2d(rev[-1], in_channels, 4, 2, 1)]
        self.decoder = nn.Sequential(*dec)
        self.lambda_weights = (1.0, 1.0, 0.25)

    def forward(self, x):
        encoding = self.encoder(x)
        q, commit, embed, idx = self.vq_layer(encoding)
        recons = self.decoder(q)
        lr, le, lc = self.lambda_weights
        losses = [lr * F.mse_loss(recons, x), le * embed, lc * commit]      # row order of J (vq_vae.py:185)
        return encoding, losses, idx


class VAEShell(nn.Module):
    """Layer shapes of the reference's `VAE` at BASELINE configs[0] (models/vae.py:28-170: 5 x (conv3x3 s2 + BatchNorm +
    LeakyReLU), hidden [32,64,128,256,512], latent 128 -> 1,701,888 parameters behind the features mu / log_var)."""

    def __init__(self, in_channels=3, hidden=(32, 64, 128, 256, 512), latent=128, size=32):
        super().__init__()
        enc, c = [], in_channels
        for h in hidden:
            enc += [nn.Conv2d(c, h, 3, 2, 1), nn.BatchNorm2d(h), nn.LeakyReLU()]
            c = h
        sp = size // 2 ** len(hidden)
        self.encoder = nn.Sequential(*enc, nn.Flatten())
        self.mu = nn.Linear(c * sp * sp, latent)
        self.log_var = nn.Linear(c * sp * sp, latent)
        self.decoder_input = nn.Linear(latent, c * sp * sp)
        rev = list(hidden)[::-1]
        dec = [nn.Unflatten(1, (c, sp, sp))]
        for a, b in zip(rev[:-1], rev[1:]):
            dec += [nn.ConvTranspose2d(a, b, 3, 2, 1, output_padding=1), nn.BatchNorm2d(b), nn.LeakyReLU()]
        dec += [nn.ConvTranspose2d(rev[-1], rev[-1], 3, 2, 1, output_padding=1), nn.BatchNorm2d(rev[-1]), nn.LeakyReLU(),
                nn.Conv2d(rev[-1], in_channels, 3, padding=1)]
        self.decoder = nn.Sequential(*dec)
        self.lambda_weights = (1.0, 0.00025)

    def forward(self, x):
        h = self.encoder(x)
        mu, log_var = self.mu(h), self.log_var(h)
        zlat = mu + torch.randn_like(mu) * torch.exp(0.5 * log_var)
        recons = self.decoder(self.decoder_input(zlat))
        kld = (-0.5 * torch.sum(1 + log_var - mu.pow(2) - log_var.exp(), dim=1)).mean()
        lr, lk = self.lambda_weights
        return [mu, log_var], [lr * F.mse_loss(recons, x), lk * kld]          # features, losses [reconstruction, kld] (vae.py:49-51)


class TorchQuantizer(nn.Module):
    """Context arm only: the quantizer written with the reference's torch expressions (vq_vae.py:27-64)."""

    def __init__(self, K, D):
        super().__init__()
        self.K, self.D = K, D
        self.embedding = nn.Embedding(K, D)
        self.embedding.weight.data.uniform_(-1 / K, 1 / K)

    def forward(self, latents):
        lat = latents.permute(0, 2, 3, 1).contiguous()
        flat = lat.view(-1, self.D)
        w = self.embedding.weight
        dist = torch.sum(flat ** 2, dim=1, keepdim=True) + torch.sum(w ** 2, dim=1) - 2 * torch.matmul(flat, w.t())
        inds = torch.argmin(dist, dim=1).unsqueeze(1)
        onehot = torch.zeros(inds.size(0), self.K, device=latents.device)
        onehot.scatter_(1, inds, 1)
        q = torch.matmul(onehot, w).view(lat.shape)
        commit = F.mse_loss(q.detach(), lat)
        embed = F.mse_loss(q, lat.detach())
        q = lat + (q - lat).detach()
        return q.permute(0, 3, 1, 2).contiguous(), commit, embed, inds.squeeze(1)


def time_train_steps(dev, batch=128, size=32, steps=20, warmup=5, agg_name="aligned_mtl"):
    """steps/s of (a) the product path: movae_b200 quantizer + mtl_backward + aggregator, and (b) for context the
    same shell with the reference's torch quantizer expressions and plain `total_loss.backward()` (the `sum` path,
    main.py:176-177: no Jacobian at all -- a lower bound on any aggregator's cost)."""
    import movae_b200

    torch.manual_seed(42)
    x = torch.rand(batch, 3, size, size, device=dev) * 2 - 1
    out = {}
    for arm in ("movae", "torch_sum"):
        torch.manual_seed(42)
        if arm == "movae":
            net = VQVAEShell(movae_b200.VectorQuantizer(512, 64)).to(dev)
            agg = movae_b200.make_aggregator(agg_name)
        else:
            net = VQVAEShell(TorchQuantizer(512, 64)).to(dev)
            agg = None
        opt = torch.optim.Adam(net.parameters(), lr=1e-4)

        def step():
            opt.zero_grad()
            encoding, losses, _ = net(x)
            if agg is None:
                sum(losses).backward()
            else:
                movae_b200.mtl_backward(losses=losses, features=[encoding], aggregator=agg, retain_graph=True)
            opt.step()

        for _ in range(warmup):
            step()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            step()
        b.record()
        torch.cuda.synchronize(dev)
        out[arm] = {"steps_per_s": round(steps / (a.elapsed_time(b) * 1e-3), 2), "ms_per_step": round(a.elapsed_time(b) / steps, 3)}
        if arm == "movae":
            shared = sum(p.numel() for p in net.encoder.parameters())
            out[arm].update({"aggregator": agg_name, "k": 3, "P_shared": shared, "N_codes": batch * (size // 4) ** 2})
    return out


def time_vae_train_steps(dev, batch=128, size=32, steps=20, warmup=5, agg_name="upgrad"):
    """BASELINE configs[0]: VAE CIFAR-10 32x32, latent 128, agg = upgrad (k = 2, P_shared = 1,701,888), batch 128."""
    import movae_b200

    torch.manual_seed(42)
    x = torch.rand(batch, 3, size, size, device=dev) * 2 - 1
    net = VAEShell().to(dev)
    agg = movae_b200.make_aggregator(agg_name)
    opt = torch.optim.Adam(net.parameters(), lr=1e-4)

    def step():
        opt.zero_grad()
        feats, losses = net(x)
        movae_b200.mtl_backward(losses=losses, features=feats, aggregator=agg, retain_graph=True)
        opt.step()

    for _ in range(warmup):
        step()
    torch.cuda.synchronize(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record()
    torch.cuda.synchronize(dev)
    shared = sum(p.numel() for m in (net.encoder, net.mu, net.log_var) for p in m.parameters())
    return {"steps_per_s": round(steps / (a.elapsed_time(b) * 1e-3), 2), "ms_per_step": round(a.elapsed_time(b) / steps, 3),
            "aggregator": agg_name, "k": 2, "P_shared": shared}
