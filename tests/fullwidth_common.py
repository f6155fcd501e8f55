"""Full-width (BASELINE config 2 widths, 256 px) evaluation of the product G / D against the oracle on the GPU -- shared by
tests/test_gpu_fullwidth.py and scripts/noise_study.py.

Every tensor (image, logits, losses, R1, all parameter gradients of the G loss, the D loss and the R1 penalty) is compared with
the oracle evaluated in fp64 (the truth).  The reference's OWN fp32 arithmetic is evaluated several times beside it -- NCHW and
channels_last tensors (different cuDNN kernels) and inputs nudged by one part in 2^22 (an fp32 rounding of the inputs) -- and
the spread of those results around fp64 is the noise level no fp32 implementation can be asked to beat: deep-layer gradients are
decided by the leaky-ReLU signs of near-zero pre-activations, and every evaluation order flips a different handful of them.
"""
import torch

DEV = 'cuda'


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def klass(k):
    return (k.split(':')[0], 'bias' if k.endswith('bias') else 'weight')


def evaluate(B, seed=3, ref_draws=4):
    """-> (rows, floor): rows = [(name, ours-vs-fp64, [reference fp32 draws vs fp64], ours-vs-first-fp32-draw)];
    floor[(group, weight|bias)] = worst distance of any reference fp32 draw from fp64 over the class."""
    from animeface_b200 import rng
    from animeface_b200.nnutils.loss import NonSaturatingLoss, r1_regularizer
    from animeface_b200.train import TrainConfig, build_models
    from oracle import sg2_torch as T
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(seed)
    cfg = TrainConfig(batch_size=B)
    G, _, D = build_models(cfg, DEV)
    sd_g = {k: v.detach().clone().requires_grad_(v.is_floating_point() and not k.endswith('.kernel')) for k, v in G.state_dict().items()}
    sd_d = {k: v.detach().clone().requires_grad_(True) for k, v in D.state_dict().items()}
    z = torch.randn(B, cfg.style_dim, device=DEV)
    real = torch.rand(B, 3, 256, 256, device=DEV) * 2 - 1
    noise = [torch.randn(B, 1, r, r, device=DEV) for r in (8, 8, 16, 16, 32, 32, 64, 64, 128, 128, 256, 256)]
    loss = NonSaturatingLoss()
    g_names = [n for n, _ in G.named_parameters()]
    d_names = [n for n, _ in D.named_parameters()]
    # ---- product path
    with rng.replay([n.clone() for n in noise]) as q:
        img, style = G(z)
        assert q.remaining == 0
    lf, lr = D(img), D(real)
    g_loss = loss.g_loss(lf)
    gg = torch.autograd.grad(g_loss, [p for p in G.parameters()], retain_graph=True, allow_unused=True)
    d_loss = loss.d_loss(lr, D(img.detach()))
    dg = torch.autograd.grad(d_loss, list(D.parameters()), allow_unused=True)
    r1 = r1_regularizer()(real, D, None)
    r1g = torch.autograd.grad(r1, list(D.parameters()), allow_unused=True)
    ours = dict(image=img, style=style, logits_fake=lf, logits_real=lr, g_loss=g_loss, d_loss=d_loss, r1=r1)
    ours.update({'ggrad:' + n: v for n, v in zip(g_names, gg)})
    ours.update({'dgrad:' + n: v for n, v in zip(d_names, dg)})
    ours.update({'r1grad:' + n: v for n, v in zip(d_names, r1g)})
    ours = {k: (None if v is None else v.detach()) for k, v in ours.items()}

    def oracle(dtype, channels_last=False, nudge=0):
        gen = torch.Generator(device=DEV).manual_seed(1000 + nudge)

        def cast(t, inp=False):
            t = t.detach().to(dtype) if t.is_floating_point() else t.detach()
            if inp and nudge:
                t = t * (1 + (torch.randint(0, 2, t.shape, device=DEV, generator=gen).to(dtype) * 2 - 1) * 2.0 ** -22)
            return t.contiguous(memory_format=torch.channels_last) if (channels_last and t.ndim == 4) else t
        g = {k: cast(v).requires_grad_(v.requires_grad) for k, v in sd_g.items()}
        d = {k: cast(v).requires_grad_(True) for k, v in sd_d.items()}
        zz, rr = cast(z, True), cast(real, True)
        o_img, o_style = T.generator(g, zz, T.ReplayDraws(T.Draws([cast(n, True) for n in noise])))
        o_lf, o_lr = T.discriminator(d, o_img, cfg.mbsd_groups), T.discriminator(d, rr, cfg.mbsd_groups)
        o_g_loss = T.g_loss_ns(o_lf)
        o_gg = torch.autograd.grad(o_g_loss, [g[n] for n in g_names], retain_graph=True, allow_unused=True)
        o_d_loss = T.d_loss_ns(o_lr, T.discriminator(d, o_img.detach(), cfg.mbsd_groups))
        o_dg = torch.autograd.grad(o_d_loss, [d[n] for n in d_names], allow_unused=True)
        o_r1 = T.r1_penalty(d, rr, cfg.mbsd_groups)
        o_r1g = torch.autograd.grad(o_r1, [d[n] for n in d_names], allow_unused=True)
        out = dict(image=o_img, style=o_style, logits_fake=o_lf, logits_real=o_lr, g_loss=o_g_loss, d_loss=o_d_loss, r1=o_r1)
        out.update({'ggrad:' + n: v for n, v in zip(g_names, o_gg)})
        out.update({'dgrad:' + n: v for n, v in zip(d_names, o_dg)})
        out.update({'r1grad:' + n: v for n, v in zip(d_names, o_r1g)})
        return {k: (None if v is None else v.detach()) for k, v in out.items()}

    o64 = oracle(torch.float64)
    variants = [(False, 0), (True, 0), (False, 1), (True, 2), (False, 3), (True, 4)][:max(2, ref_draws)]
    refs = [oracle(torch.float32, cl, nd) for cl, nd in variants]
    rows, floor = [], {}
    for k, truth in o64.items():
        a = ours[k]
        if truth is None:
            assert a is None or float(a.abs().max()) == 0, k
            continue
        if float(truth.abs().max()) < 1e-12:
            continue
        assert a is not None, k
        draws = [rel(r[k], truth) for r in refs]
        rows.append((k, rel(a, truth), draws, rel(a, refs[0][k])))
        floor[klass(k)] = max(floor.get(klass(k), 0.0), max(draws))
    return rows, floor


GROUPS = ('image', 'style', 'logits_fake', 'logits_real', 'g_loss', 'd_loss', 'r1', 'ggrad', 'dgrad', 'r1grad')


def report(B, rows, floor, bar=1e-3):
    lines = [f'full-width parity at B = {B}: vs fp64 (ours | worst / median of the reference fp32 draws) and ours vs the first fp32 draw; worst per group:']
    for grp in GROUPS:
        sel = [r for r in rows if r[0].split(':')[0] == grp]
        if not sel:
            continue
        k, e, draws, _ = max(sel, key=lambda r: r[1])
        worst_ref = max(max(r[2]) for r in sel)
        med = sorted(r[1] for r in sel)[len(sel) // 2]
        med_ref = sorted(sorted(r[2])[len(r[2]) // 2] for r in sel)[len(sel) // 2]
        k2, _, _, d32 = max(sel, key=lambda r: r[3])
        lines.append(f'   {grp:12s} ours-fp64 worst {e:.2e} median {med:.2e} | reference fp32-fp64 worst {worst_ref:.2e} median {med_ref:.2e}  '
                     f'({sum(1 for r in sel if r[1] > bar)}/{len(sel)} of ours above {bar:g}; worst: {k})  || ours-fp32 {d32:.2e}')
    lines.append('   fp32 noise level of the reference per class (worst draw vs fp64): '
                 + ', '.join(f'{a}/{b} {v:.1e}' for (a, b), v in sorted(floor.items()) if a.endswith('grad')))
    return '\n'.join(lines)
