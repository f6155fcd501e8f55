import ast, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from animeface_b200.model import Discriminator, Generator, supplied_noise
from animeface_b200.nnutils.loss import NonSaturatingLoss
from animeface_b200.ops import conv2d as C
g = np.load(os.path.join(ROOT, 'tests/golden/model.npz'))
c = ast.literal_eval(str(g['cfg']))
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
def rel(a, b):
    b = np.asarray(b, np.float64); a = a.detach().double().cpu().numpy()
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
for impl in (1, 0):
    C.set_default_impl(impl)
    G = Generator(c['image_size'], 3, c['style_dim'], c['channels'], c['max_channels'], 2, c['map_num_layers'], True, 0.01).cuda()
    D = Discriminator(c['image_size'], 3, c['channels'], c['max_channels'], 2, c['mbsd_groups']).cuda()
    G.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('G0.')})
    D.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('D0.')})
    z, real = T(g['z']), T(g['real'])
    with supplied_noise([T(g[f'fwd.noise.{i}']) for i in range(int(g['fwd.n_noise']))]):
        image, style = G(z)
    lf, lr = D(image.detach()), D(real)
    d_loss = NonSaturatingLoss().d_loss(lr, lf)
    names = [n for n, _ in D.named_parameters()]
    dg = torch.autograd.grad(d_loss, list(D.parameters()), allow_unused=True)
    print('impl', impl, 'image', rel(image, g['fwd.image']), 'lr', rel(lr, g['fwd.logits_real']), 'd_loss', float(d_loss), float(g['d_loss']))
    for n, gr in zip(names, dg):
        print(f'   {n:36s} {rel(gr, g["dgrad." + n]):.2e}')

print('---- bisect: which op class on the tensor cores causes it')
orig_conv, orig_wgrad = C._conv_raw, C._wgrad_raw
for label, fwd_impl, dg_impl, wg_impl in (('fwd only', 0, 1, 1), ('dgrad only', 1, 0, 1), ('wgrad only', 1, 1, 0), ('fwd TC no refine', 0, 1, 1)):
    def conv(x, w, coef, transpose, *a, **k):
        k['impl'] = None
        C._default_impl = dg_impl if transpose else fwd_impl
        if label.endswith('no refine'):
            k['refine'] = False
        return orig_conv(x, w, coef, transpose, *a, **k)
    def wgrad(x, gy, kk, coef, **k):
        k['impl'] = wg_impl
        return orig_wgrad(x, gy, kk, coef, **k)
    C._conv_raw, C._wgrad_raw = conv, wgrad
    D = Discriminator(c['image_size'], 3, c['channels'], c['max_channels'], 2, c['mbsd_groups']).cuda()
    D.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('D0.')})
    lf, lr = D(T(g['fwd.image'])), D(T(g['real']))
    d_loss = NonSaturatingLoss().d_loss(lr, lf)
    names = [n for n, _ in D.named_parameters()]
    dg = torch.autograd.grad(d_loss, list(D.parameters()), allow_unused=True)
    errs = {n: rel(gr, g['dgrad.' + n]) for n, gr in zip(names, dg)}
    worst = max(errs, key=errs.get)
    print(f'{label:18s} worst {worst} {errs[worst]:.2e}   blocks.1.block.0.w {errs["blocks.1.block.0.layer.weight"]:.2e}  lr err {rel(lr, g["fwd.logits_real"]):.1e}')
