"""Bench plumbing (NOT part of the product path): a VQ-VAE model shell with the layer shapes of the
reference's `VQVAE` at the BASELINE config (models/vq_vae.py:148-318 with hidden_dims [128, 256],
2 residual layers, K = 512, D = 64: 2,448,064 shared encoder parameters, 1,988,995 decoder parameters),
written here because /root/reference does not exist on the GPU box.  Convolutions are plain torch.nn
(cuDNN); the quantizer and the multi-objective backward are the product's.

train_step() mirrors the reference's train_epoch body (main.py:157-214): zero_grad -> forward -> loss dict
[reconstruction, embedding, commitment] (vq_vae.py:185, :380-390, lambda 1 / 1 / 0.25) ->
mtl_backward(losses, features=[encoding], aggregator) -> Adam step.

Also here: the GG-VQ-VAE v1 objective set (gg_vq_vae.py:63-66, :151-164: k = 4 with the Sobel-weighted pixel loss)
on the same network (BASELINE configs[2]) and a VQ-VAE-2 shell (vq_vae2.py:31-233: bottom/top encoders, two
quantizers, features [encoding_top, encoding_bottom], loss order [reconstruction, commitment, embedding];
BASELINE configs[3]).

Arms timed per config (`time_config`):
  movae_eager      product path, eager launches, torch.optim.Adam (per-tensor optimizer)
  movae_graph      product path, FlatParameters + fused Adam (K7), whole step replayed from ONE CUDA graph
  movae_graph_e2e  the same graph, plus per step the H2D copy of the image batch from pinned host memory and the D2H
                   read of the loss vector (what a training loop around it pays)
  torch_sum_*      context only: the reference's torch quantizer expressions and `total_loss.backward()` (the `sum`
                   path, main.py:176-177: no Jacobian at all -- a lower bound on any aggregator's cost), eager / graphed
  reference_style  context only, the SAME algorithm as the product arm run the way the reference runs it on a GPU
                   (`reference_style_step`): torch quantizer, torchjd-style mtl_backward (one vmapped backward pass,
                   per-parameter reshape + `cat`), `J @ J.T`, the weighting with the reference's own torch expressions
                   (UPGrad: `G.cpu()` -> float64 QP on the host -> device), `w @ J`, per-tensor `clone` into `.grad`,
                   the two hooks of main.py:71-122 with their `.item()`s, torch Adam.  This is what the product replaces.
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


class GGVQVAEShell(VQVAEShell):
    """GG-VQ-VAE v1 (gg_vq_vae.py:13-164): the VQ-VAE network with a fourth objective, the Sobel-edge-weighted pixel
    loss `gradient_guided_loss`; row order [reconstruction, embedding, commitment, gradient_guided], lambda 1/1/0.25/1."""

    def __init__(self, quantizer: nn.Module, **kw):
        super().__init__(quantizer, **kw)
        sx = torch.tensor([[-1., 0., 1.], [-2., 0., 2.], [-1., 0., 1.]]).view(1, 1, 3, 3)
        sy = torch.tensor([[-1., -2., -1.], [0., 0., 0.], [1., 2., 1.]]).view(1, 1, 3, 3)
        self.register_buffer("sobel_x", sx.expand(3, 1, 3, 3).clone())
        self.register_buffer("sobel_y", sy.expand(3, 1, 3, 3).clone())

    def forward(self, x):
        encoding = self.encoder(x)
        q, commit, embed, idx = self.vq_layer(encoding)
        recons = self.decoder(q)
        gx = F.conv2d(x, self.sobel_x, padding=1, groups=3)
        gy = F.conv2d(x, self.sobel_y, padding=1, groups=3)
        wgt = torch.sqrt(gx ** 2 + gy ** 2 + 1e-8).max(dim=1)[0]
        wgt = wgt / (wgt.max() + 1e-8)
        gg = (wgt.unsqueeze(1) * F.mse_loss(recons, x, reduction="none")).mean()
        losses = [F.mse_loss(recons, x), embed, 0.25 * commit, gg]
        return encoding, losses, idx


class _Res2(nn.Module):
    def __init__(self, c: int, r: int):
        super().__init__()
        self.conv = nn.Sequential(nn.ReLU(), nn.Conv2d(c, r, 3, padding=1), nn.ReLU(), nn.Conv2d(r, c, 1))

    def forward(self, x):
        return x + self.conv(x)


class VQVAE2Shell(nn.Module):
    """Layer shapes of the reference's `VQVAE2` (vq_vae2.py:31-233) with hidden_dims[0] = 128, 2 residual blocks of 32
    channels, K = 512, D = 64: 651,392 parameters behind the features (enc_b + enc_t)."""

    def __init__(self, make_quantizer, in_channels=3, c=128, n_res=2, r=32, D=64, K=512):
        super().__init__()
        res = lambda: [_Res2(c, r) for _ in range(n_res)] + [nn.ReLU()]  # noqa: E731
        self.enc_b = nn.Sequential(nn.Conv2d(in_channels, c // 2, 4, 2, 1), nn.ReLU(), nn.Conv2d(c // 2, c, 4, 2, 1), nn.ReLU(),
                                   nn.Conv2d(c, c, 3, padding=1), *res())
        self.enc_t = nn.Sequential(nn.Conv2d(c, c // 2, 4, 2, 1), nn.ReLU(), nn.Conv2d(c // 2, c, 3, padding=1), *res())
        self.quantize_conv_t = nn.Conv2d(c, D, 1)
        self.quantize_t = make_quantizer(K, D)
        self.dec_t = nn.Sequential(nn.Conv2d(D, c, 3, padding=1), *res(), nn.ConvTranspose2d(c, D, 4, 2, 1))
        self.quantize_conv_b = nn.Conv2d(D + c, D, 1)
        self.quantize_b = make_quantizer(K, D)
        self.upsample_t = nn.ConvTranspose2d(D, D, 4, 2, 1)
        self.dec = nn.Sequential(nn.Conv2d(2 * D, c, 3, padding=1), *res(), nn.ConvTranspose2d(c, c // 2, 4, 2, 1), nn.ReLU(),
                                 nn.ConvTranspose2d(c // 2, in_channels, 4, 2, 1))

    def forward(self, x):
        enc_b = self.enc_b(x)
        enc_t = self.enc_t(enc_b)
        qt, commit_t, embed_t, idx_t = self.quantize_t(self.quantize_conv_t(enc_t))
        dec_t = self.dec_t(qt)
        qb, commit_b, embed_b, idx_b = self.quantize_b(self.quantize_conv_b(torch.cat([dec_t, enc_b], 1)))
        recons = self.dec(torch.cat([self.upsample_t(qt), qb], 1))
        losses = [F.mse_loss(recons, x), commit_t + commit_b, embed_t + embed_b]   # vq_vae2.py:141-145 order, lambda 1/1/1
        return [enc_t, enc_b], losses, (idx_t, idx_b)


# ---------------------------------------------------------------------------------------------- reference-style arm
def _ref_style_weights(name: str, G: torch.Tensor, losses):
    """The weighting as the reference computes it, on G's device (context arm; restates aligned_mtl.py:97-133 and
    mgda.py:221-285,:343-367 with their torch expressions and host syncs; UPGrad goes through the host like torchjd)."""
    k = G.shape[0]
    dev = G.device
    if name == "upgrad":
        from oracle import aggregation as oa
        return oa.upgrad_weights(G.cpu()).to(dev)
    if name.startswith("aligned_mtl"):
        w0 = torch.full((k,), 1.0 / k, device=dev)
        lam, V = torch.linalg.eigh(G, UPLO="U")
        tol = torch.max(lam) * k * torch.finfo().eps
        rank = int(sum(lam > tol))                                  # aligned_mtl.py:110: python sum over a tensor
        if rank == 0:
            return w0
        order = torch.argsort(lam, dim=-1, descending=True)
        lam, V = lam[order][:rank], V[:, order][:, :rank]
        scale = lam[-1] if name == "aligned_mtl" else torch.median(lam)
        B = scale.sqrt() * V @ torch.diag(1 / lam.sqrt()) @ V.T
        return B @ w0
    if name.startswith("mgda"):
        ell = losses.detach().clamp(min=1e-20)
        nrm = torch.sqrt(torch.diag(G).clamp(min=1e-20))
        s = {"mgda_ln": nrm, "mgda_gn": ell, "mgda_lgn": ell * nrm}.get(name)
        R = G if s is None else G / (s.unsqueeze(1) * s.unsqueeze(0))
        alpha = torch.ones(k, device=dev) / k
        for _ in range(250):
            t = torch.argmin(R @ alpha)
            e_t = torch.zeros(k, device=dev)
            e_t[t] = 1.0
            a, b, c = alpha @ (R @ e_t), alpha @ (R @ alpha), e_t @ (R @ e_t)
            if c <= a:
                gamma = 1.0
            elif b <= a:
                gamma = 0.0
            else:
                gamma = (b - a) / (b + c - 2 * a)
            alpha = (1 - gamma) * alpha + gamma * e_t
            if gamma < 1e-5:
                break
        return alpha
    raise ValueError(name)


def reference_style_step(net, xs, agg_name: str, opt, state: dict):
    """One train step the way the reference runs it (main.py:157-214 with torchjd's mtl_backward)."""
    from movae_b200.autojac import _leaves_of

    opt.zero_grad()
    feats, losses, _ = net(xs)
    feats = feats if isinstance(feats, list) else [feats]
    if "shared" not in state:
        state["shared"] = _leaves_of(feats)
        ids = {id(p) for p in state["shared"]}
        state["tasks"] = [[p for p in _leaves_of([l], stop_at=feats) if id(p) not in ids] for l in losses]
    shared, tasks = state["shared"], state["tasks"]
    k = len(losses)
    loss_vec = torch.stack([l.detach() for l in losses])
    feat_grads = []
    for loss, tparams in zip(losses, tasks):
        outs = torch.autograd.grad(loss, feats + tparams, retain_graph=True, allow_unused=True)
        feat_grads.append([torch.zeros_like(f) if g is None else g for f, g in zip(feats, outs[:len(feats)])])
        for p, g in zip(tparams, outs[len(feats):]):
            if g is not None:
                p.grad = g.clone() if p.grad is None else p.grad + g
    stacked = [torch.stack([fg[j] for fg in feat_grads]) for j in range(len(feats))]
    jac = torch.autograd.grad(feats, shared, grad_outputs=stacked, is_grads_batched=True, allow_unused=True)
    J = torch.cat([(torch.zeros(k, p.numel(), device=xs.device) if j is None else j.reshape(k, -1)) for p, j in zip(shared, jac)], dim=1)
    G = J @ J.T                                                    # torchjd compute_gramian
    w = _ref_style_weights(agg_name, G, loss_vec)
    state["weights"] = [w[i].item() for i in range(k)]             # hook print_weights, main.py:71-91
    g = w @ J                                                      # WeightedAggregator.forward
    state["similarity"] = F.cosine_similarity(g, J.mean(dim=0), dim=0).item()   # hook print_gd_similarity, main.py:94-122
    off = 0
    for p in shared:                                               # split / _Reshape / Accumulate
        n = p.numel()
        p.grad = g[off:off + n].view(p.shape).clone()
        off += n
    opt.step()
    return loss_vec


def reference_style_runner(kind: str, dev):
    """A ready-to-call reference-style step for BASELINE configs[0] ("vae", agg upgrad) or configs[1] ("vqvae",
    aligned_mtl) on `dev` (also the CPU: bench.py's cpu_baseline leg)."""
    torch.manual_seed(42)
    if kind == "vae":
        class _VAE(VAEShell):
            def forward(self, x):
                feats, losses = super().forward(x)
                return feats, losses, None
        net, agg_name, batch, size = _VAE().to(dev), "upgrad", 128, 32
    else:
        net, agg_name, batch, size = VQVAEShell(TorchQuantizer(512, 64)).to(dev), "aligned_mtl", 128, 32
    xs = torch.rand(batch, 3, size, size, device=dev) * 2 - 1
    opt = torch.optim.Adam(net.parameters(), lr=1e-4)
    state: dict = {}
    return lambda: reference_style_step(net, xs, agg_name, opt, state)


def _timed(step, steps: int, warmup: int, dev):
    for _ in range(warmup):
        step()
    torch.cuda.synchronize(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record()
    torch.cuda.synchronize(dev)
    ms = a.elapsed_time(b) / steps
    return {"steps_per_s": round(1e3 / ms, 2), "ms_per_step": round(ms, 3)}


def time_config(dev, build, batch: int, size: int, agg_name: str, steps: int = 20, warmup: int = 5,
                arms=("movae_eager", "movae_graph", "movae_graph_e2e", "torch_sum_eager", "torch_sum_graph", "reference_style")):
    """`build(make_quantizer) -> net` whose forward returns (features, losses, indices).  See the module docstring
    for the arms.  Synthetic images `rand * 2 - 1`, seed 42."""
    import time

    import movae_b200

    out = {}
    torch.manual_seed(42)
    x = torch.rand(batch, 3, size, size, device=dev) * 2 - 1
    h_x = x.cpu().pin_memory()
    info = {}
    for arm in arms:
        torch.manual_seed(42)
        product = arm.startswith("movae")
        net = build(movae_b200.VectorQuantizer if product else TorchQuantizer).to(dev)
        agg = movae_b200.make_aggregator(agg_name) if product else None
        graph = "graph" in arm
        if product and graph:
            opt = movae_b200.Adam(net.parameters(), lr=1e-4)
        else:
            opt = torch.optim.Adam(net.parameters(), lr=1e-4, capturable=graph)
        xs = x.clone()
        if arm == "reference_style":
            state: dict = {}
            out[arm] = _timed(lambda: reference_style_step(net, xs, agg_name, opt, state), steps, warmup, dev)
            del net, opt
            continue

        def step():
            opt.zero_grad()
            feats, losses, _ = net(xs)
            if agg is None:
                sum(losses).backward()
            else:
                if isinstance(agg, movae_b200.MGDA):
                    agg.set_losses(torch.stack([l.detach() for l in losses]))          # main.py:185-186
                movae_b200.mtl_backward(losses=losses, features=feats if isinstance(feats, list) else [feats],
                                        aggregator=agg, retain_graph=True)
            opt.step()
            return torch.stack([l.detach() for l in losses])

        if graph:
            g = movae_b200.GraphedStep(step, warmup=3)
            if arm.endswith("e2e"):
                def run():
                    xs.copy_(h_x, non_blocking=True)          # H2D of the batch, every step
                    return g().cpu()                           # D2H of the k losses (synchronises, like main.py:216-218)
                for _ in range(warmup):
                    run()
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                for _ in range(steps):
                    run()
                torch.cuda.synchronize(dev)
                ms = 1e3 * (time.perf_counter() - t0) / steps
                out[arm] = {"steps_per_s": round(1e3 / ms, 2), "ms_per_step": round(ms, 3),
                            "h2d_bytes_per_step": h_x.numel() * 4, "d2h_bytes_per_step": 4 * len(g.outputs)}
            else:
                out[arm] = _timed(g, steps, warmup, dev)
        else:
            out[arm] = _timed(step, steps, warmup, dev)
        if product and not info:
            feats, losses, _ = net(xs)
            feats = feats if isinstance(feats, list) else [feats]
            from movae_b200.autojac import _leaves_of
            info = {"aggregator": agg_name, "k": len(losses), "P_shared": sum(p.numel() for p in _leaves_of(feats)),
                    "P_total": sum(p.numel() for p in net.parameters())}
        del net, opt
    out.update(info)
    return out


def time_train_steps(dev, batch=128, size=32, steps=20, warmup=5, agg_name="aligned_mtl"):
    """BASELINE configs[1]: VQ-VAE CIFAR-10 32x32, K=512, D=64, hidden [128,256], agg = aligned_mtl, batch 128."""
    r = time_config(dev, lambda mq: VQVAEShell(mq(512, 64)), batch, size, agg_name, steps, warmup)
    r["N_codes"] = batch * (size // 4) ** 2
    return r


def time_ggvqvae_train_steps(dev, batch=256, size=64, steps=10, warmup=3, agg_name="mgda_lgn"):
    """BASELINE configs[2]: GG-VQ-VAE CelebA 64x64, agg = mgda_lgn, batch 256 (k = 4, N = 65,536 code vectors)."""
    r = time_config(dev, lambda mq: GGVQVAEShell(mq(512, 64)), batch, size, agg_name, steps, warmup)
    r["N_codes"] = batch * (size // 4) ** 2
    return r


def time_vqvae2_train_steps(dev, batch=64, size=256, steps=5, warmup=2, agg_name="upgrad"):
    """BASELINE configs[3]: VQ-VAE2 CelebA-HQ 256x256 (top + bottom codebooks), agg = upgrad, batch 64."""
    r = time_config(dev, lambda mq: VQVAE2Shell(mq), batch, size, agg_name, steps, warmup,
                    arms=("movae_eager", "movae_graph", "movae_graph_e2e", "torch_sum_graph", "reference_style"))
    r["N_codes"] = {"top": batch * (size // 8) ** 2, "bottom": batch * (size // 4) ** 2}
    return r


def time_vae_train_steps(dev, batch=128, size=32, steps=20, warmup=5, agg_name="upgrad"):
    """BASELINE configs[0]: VAE CIFAR-10 32x32, latent 128, agg = upgrad (k = 2, P_shared = 1,701,888), batch 128."""
    class _VAE(VAEShell):
        def forward(self, x):
            feats, losses = super().forward(x)
            return feats, losses, None

    return time_config(dev, lambda mq: _VAE(), batch, size, agg_name, steps, warmup,
                       arms=("movae_eager", "movae_graph", "movae_graph_e2e", "torch_sum_graph", "reference_style"))


def time_dp_train_steps(dev, rank: int, world: int, batch=128, size=32, steps=20, warmup=5, agg_name="aligned_mtl"):
    """Data-parallel train step (movae_b200.parallel.DataParallel) of BASELINE configs[1] at `world` GPUs, weak scaling:
    every rank holds a replica and its own batch of `batch` images; Jacobian rows reduce-scattered, K1 / K3 on the column
    shard, Gramian all_reduce, aggregated gradient all-gathered, decoder / codebook gradients all-reduced.  Arms: eager
    launches, and the whole step (collectives included) replayed from one CUDA graph.  Collective call on all ranks; timing
    = max over ranks between barriers.  Returns the dict on every rank."""
    import torch.distributed as dist

    import movae_b200
    from movae_b200 import parallel

    out = {}
    for arm in ("eager", "graph"):
        torch.manual_seed(42)                                         # identical replicas
        net = VQVAEShell(movae_b200.VectorQuantizer(512, 64)).to(dev)
        gen = torch.Generator(device=dev).manual_seed(1000 + rank)    # a different batch per rank
        xs = torch.rand(batch, 3, size, size, generator=gen, device=dev) * 2 - 1
        agg = movae_b200.make_aggregator(agg_name)
        parallel.DataParallel(agg)
        opt = movae_b200.Adam(net.parameters(), lr=1e-4)

        def step():
            opt.zero_grad()
            feats, losses, _ = net(xs)
            movae_b200.mtl_backward(losses=losses, features=[feats], aggregator=agg, retain_graph=True)
            opt.step()
            return torch.stack([l.detach() for l in losses])

        fn = movae_b200.GraphedStep(step, warmup=3) if arm == "graph" else step
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize(dev)
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        torch.cuda.synchronize(dev)
        dist.barrier()
        t = torch.tensor([a.elapsed_time(b) / steps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        # replicas must still agree after the timed steps
        chk = next(net.parameters()).detach().double().sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        out[arm] = {"steps_per_s": round(1e3 / ms, 2), "ms_per_step": round(ms, 3), "images_per_s": round(world * batch * 1e3 / ms, 1),
                    "replicas_identical": bool(float(lo) == float(hi))}
        del fn, opt, net
    out.update({"aggregator": agg_name, "k": 3, "global_batch": world * batch, "per_gpu_batch": batch, "world": world,
                "scaling": "weak (per-GPU batch fixed)"})
    return out
