"""GPU parity at the FULL channel widths of BASELINE config 2 (channels 32..512, 256 px): the product Generator /
Discriminator against the oracle (oracle/sg2_torch.py, plain torch fp32 with TF32 off) evaluated on the same GPU with the
same weights, latents and noise.  The reference-generated goldens pin the oracle on a small model (test_oracle_golden.py);
this test carries that pin to the layer shapes the tcgen05 kernels actually run at (ci, co in 32..512, k = 1 and 3,
4^2..256^2), which the small golden model cannot reach.  Bar: north_star's 1e-3 relative to each tensor's scale."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda'
BAR = 1e-3


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize('B', [4, 32])
def test_full_width_model_vs_oracle_on_gpu(B):
    """B = 32 is BASELINE config 2's batch (one forward + backward of everything fits in the 180 GB); B = 4 is the quick case."""
    from animeface_b200 import rng
    from animeface_b200.nnutils.loss import NonSaturatingLoss, r1_regularizer
    from animeface_b200.train import TrainConfig, build_models
    from oracle import sg2_torch as T
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(3)
    cfg = TrainConfig(batch_size=B)
    G, _, D = build_models(cfg, DEV)
    sd_g = {k: v.detach().clone().requires_grad_(v.is_floating_point() and not k.endswith('.kernel')) for k, v in G.state_dict().items()}
    sd_d = {k: v.detach().clone().requires_grad_(True) for k, v in D.state_dict().items()}
    z = torch.randn(B, cfg.style_dim, device=DEV)
    real = torch.rand(B, 3, 256, 256, device=DEV) * 2 - 1
    noise = [torch.randn(B, 1, r, r, device=DEV) for r in (8, 8, 16, 16, 32, 32, 64, 64, 128, 128, 256, 256)]
    loss = NonSaturatingLoss()
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
    # ---- oracle on the same device: fp32 (the reference's arithmetic) and fp64 (the truth both are measured against)
    def oracle(dtype, channels_last=False):
        def cast(t):
            t = t.detach().to(dtype) if t.is_floating_point() else t.detach()
            return t.contiguous(memory_format=torch.channels_last) if (channels_last and t.ndim == 4) else t
        g = {k: cast(v).requires_grad_(v.requires_grad) for k, v in sd_g.items()}
        d = {k: cast(v).requires_grad_(True) for k, v in sd_d.items()}
        o_img, o_style = T.generator(g, cast(z), T.ReplayDraws(T.Draws([cast(n) for n in noise])))
        o_lf, o_lr = T.discriminator(d, o_img, cfg.mbsd_groups), T.discriminator(d, cast(real), cfg.mbsd_groups)
        o_g_loss = T.g_loss_ns(o_lf)
        o_gg = torch.autograd.grad(o_g_loss, [g[n] for n in g_names], retain_graph=True, allow_unused=True)
        o_d_loss = T.d_loss_ns(o_lr, T.discriminator(d, o_img.detach(), cfg.mbsd_groups))
        o_dg = torch.autograd.grad(o_d_loss, [d[n] for n in d_names], allow_unused=True)
        o_r1 = T.r1_penalty(d, cast(real), cfg.mbsd_groups)
        o_r1g = torch.autograd.grad(o_r1, [d[n] for n in d_names], allow_unused=True)
        out = dict(image=o_img, style=o_style, logits_fake=o_lf, logits_real=o_lr, g_loss=o_g_loss, d_loss=o_d_loss, r1=o_r1)
        out.update({'ggrad:' + n: v for n, v in zip(g_names, o_gg)})
        out.update({'dgrad:' + n: v for n, v in zip(d_names, o_dg)})
        out.update({'r1grad:' + n: v for n, v in zip(d_names, o_r1g)})
        return {k: (None if v is None else v.detach()) for k, v in out.items()}

    g_names = [n for n, _ in G.named_parameters()]
    d_names = [n for n, _ in D.named_parameters()]
    ours = dict(image=img, style=style, logits_fake=lf, logits_real=lr, g_loss=g_loss, d_loss=d_loss, r1=r1)
    ours.update({'ggrad:' + n: v for n, v in zip(g_names, gg)})
    ours.update({'dgrad:' + n: v for n, v in zip(d_names, dg)})
    ours.update({'r1grad:' + n: v for n, v in zip(d_names, r1g)})
    o32, o64 = oracle(torch.float32), oracle(torch.float64)
    o32b = oracle(torch.float32, channels_last=True)      # the reference's arithmetic once more, through other cuDNN kernels

    # Each tensor is judged against the fp64 evaluation.  The bar is 1e-3 of the tensor's scale; where the reference's own
    # fp32 arithmetic is itself further than 1e-3/3 from fp64 -- deep-layer gradients decided by leaky-ReLU signs of
    # near-zero pre-activations -- no fp32 implementation can be asked for more than the reference delivers, and the bar
    # becomes 3x the reference's own distance.  That distance is a NOISE level, not a per-tensor constant (every fp32
    # evaluation order flips a different handful of signs: scripts/r1_subnet_check.py shows the reference's fp32 result
    # 10..1000x further from fp64 than ours on some tensors and 3x closer on others), so it is estimated per class of
    # quantity -- (gradient kind, weight | bias) -- as the worst distance of two fp32 evaluations of the reference (NCHW
    # and channels_last tensors: different cuDNN kernels, same arithmetic class).
    def klass(k):
        return (k.split(':')[0], 'bias' if k.endswith('bias') else 'weight')

    floor = {}
    for k, truth in o64.items():
        if truth is None or float(truth.abs().max()) < 1e-12:
            continue
        floor[klass(k)] = max(floor.get(klass(k), 0.0), _rel(o32[k], truth), _rel(o32b[k], truth))
    rows, bad = [], []
    for k, truth in o64.items():
        a = ours[k]
        if truth is None:
            assert a is None or float(a.abs().max()) == 0, k
            continue
        if float(truth.abs().max()) < 1e-12:
            continue
        assert a is not None, k
        e, e32 = _rel(a, truth), _rel(o32[k], truth)
        rows.append((k, e, e32, _rel(a, o32[k])))
        if e > max(BAR, 3 * floor[klass(k)]):
            bad.append((k, e, e32, floor[klass(k)]))
    print(f'\nfull-width parity at B = {B}: vs fp64 (ours | reference fp32 arithmetic) and ours vs the fp32 reference arithmetic, worst per group:')
    for grp in ('image', 'style', 'logits_fake', 'logits_real', 'g_loss', 'd_loss', 'r1', 'ggrad', 'dgrad', 'r1grad'):
        sel = [r for r in rows if r[0].split(':')[0] == grp]
        if sel:
            k, e, e32, _ = max(sel, key=lambda r: r[1])
            over = sum(1 for r in sel if r[1] > BAR)
            k2, _, _, d32 = max(sel, key=lambda r: r[3])
            over2 = sum(1 for r in sel if r[3] > BAR)
            print(f'   {grp:12s} ours-fp64 {e:.2e} | fp32oracle-fp64 {e32:.2e}  worst: {k} ({over}/{len(sel)} above 1e-3)'
                  f'   || ours-fp32oracle {d32:.2e}  worst: {k2} ({over2}/{len(sel)} above 1e-3)')
    print('   fp32 noise level of the reference per class (worst of two fp32 evaluations vs fp64): '
          + ', '.join(f'{a}/{b} {v:.1e}' for (a, b), v in sorted(floor.items()) if a.endswith('grad')))
    assert not bad, bad
