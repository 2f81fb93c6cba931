"""CPU: the product's HOST code (conv2d_resample case split, modulated_conv2d, AugmentPipe, networks, loss) driven with
the oracle's primitive ops (oracle/backend.py) against golden vectors from the real reference.  This pins everything
above the kernels: shapes, gains, clamps, paddings, RNG draw order, which derivatives exist."""
import numpy as np
import pytest
import torch

from oracle import gen_golden as gg
from oracle.backend import oracle_ops
from gan_track_b200.training import augment, loss as loss_mod, networks_stylegan2 as nets
from gan_track_b200.torch_utils.ops import conv2d_resample, upfirdn2d


def t(a):
    return torch.from_numpy(np.asarray(a))


def close(a, b, rtol=1e-4, atol=1e-5):
    a = a.detach().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


@pytest.mark.parametrize('case', gg.CONV_RESAMPLE_CASES, ids=lambda c: c[0])
def test_conv2d_resample_case_split(golden, case):
    name, ci, co, k, kw, h = case
    G = golden('ops_conv.npz')
    x = t(G[f'conv2d_resample/{name}/x']).requires_grad_(True)
    w = t(G[f'conv2d_resample/{name}/w']).requires_grad_(True)
    with oracle_ops():
        y = conv2d_resample.conv2d_resample(x, w, f=upfirdn2d.setup_filter(gg.F4), **kw)
        dx, dw = torch.autograd.grad(y, [x, w], t(G[f'conv2d_resample/{name}/dy']))
    close(y, G[f'conv2d_resample/{name}/y'])
    close(dx, G[f'conv2d_resample/{name}/dx'])
    close(dw, G[f'conv2d_resample/{name}/dw'], 1e-4, 1e-4)


@pytest.mark.parametrize('name,kw', [('plain_demod_noise', dict(padding=1)), ('up_demod_noise', dict(up=2, padding=1, flip_weight=False)),
                                     ('torgb', dict(demodulate=False))])
def test_modulated_conv2d_host(golden, name, kw):
    G = golden('ops_conv.npz')
    x = t(G[f'modconv/{name}/x']).requires_grad_(True)
    w = t(G[f'modconv/{name}/w']).requires_grad_(True)
    s = t(G[f'modconv/{name}/s']).requires_grad_(True)
    noise = t(G[f'modconv/{name}/noise']) if f'modconv/{name}/noise' in G else None
    dy = t(G[f'modconv/{name}/dy'])
    for fused in (False, True):
        tag = f'modconv/{name}/{"fused" if fused else "unfused"}'
        with oracle_ops():
            y = nets.modulated_conv2d(x, w, s, noise=(noise.clone() if noise is not None else None),
                                      resample_filter=upfirdn2d.setup_filter(gg.F4), fused_modconv=fused, **kw)
            dx, dw, ds = torch.autograd.grad(y, [x, w, s], dy)
        close(y, G[f'{tag}/y'])
        close(dx, G[f'{tag}/dx'])
        close(dw, G[f'{tag}/dw'], 1e-4, 1e-4)
        close(ds, G[f'{tag}/ds'], 1e-4, 1e-4)


@pytest.mark.parametrize('tag,kw,key', [('claro', gg.AUG_KW, 'img'), ('full', gg.AUG_KW_FULL, 'img'), ('full_rgb', gg.AUG_KW_FULL, 'rgb')])
def test_augment_pipe_known_answers(golden, tag, kw, key):
    G = golden('augment.npz')
    x = t(G[f'augment/{key}'])
    pipe = augment.AugmentPipe(run_dir=None, batch_size=x.shape[0], **kw)
    pipe.p.copy_(torch.as_tensor(0.7))
    with oracle_ops():
        for pct in (0.1, 0.5, 0.9):
            torch.manual_seed(11)
            close(pipe(x, False, debug_percentile=pct), G[f'augment/{tag}/pct{pct}'], 1e-4, 2e-5)
        torch.manual_seed(123)
        close(pipe(x, False), G[f'augment/{tag}/seed123'], 1e-4, 2e-5)


def test_augment_pipe_double_backward(golden):
    G = golden('augment.npz')
    pipe = augment.AugmentPipe(**gg.AUG_KW)
    pipe.p.copy_(torch.as_tensor(1.0))
    x = t(G['augment/img']).clone().requires_grad_(True)
    with oracle_ops():
        torch.manual_seed(5)
        y = pipe(x, False)
        gx, = torch.autograd.grad((y * t(G['augment/grad/w'])).sum(), x, create_graph=True)
        ggx, = torch.autograd.grad(gx.square().sum(), x)
    close(gx, G['augment/grad/gx'], 1e-4, 2e-5)
    close(ggx, G['augment/grad/ggx'], 1e-3, 1e-4)


def _load(module, G, prefix):
    sd = {k[len(prefix):]: t(G[k]) for k in G.keys(prefix)}
    missing, unexpected = module.load_state_dict(sd, strict=True), None
    return module


@pytest.fixture(scope='module')
def models(golden):
    G = golden('model.npz')
    gen = _load(nets.Generator(**gg.G_KW), G, 'model/G/').train().requires_grad_(False)
    dis = _load(nets.Discriminator(**gg.D_KW), G, 'model/D/').train().requires_grad_(False)
    return G, gen, dis


def test_state_dict_names_match_reference(models):
    G, gen, dis = models
    assert sorted(gen.state_dict().keys()) == sorted(k[len('model/G/'):] for k in G.keys('model/G/'))
    assert sorted(dis.state_dict().keys()) == sorted(k[len('model/D/'):] for k in G.keys('model/D/'))
    for k, v in gen.state_dict().items():
        assert tuple(v.shape) == G['model/G/' + k].shape, k


def test_generator_and_discriminator_outputs(models):
    G, gen, dis = models
    z, c, real = t(G['model/z']), t(G['model/c']), t(G['model/real'])
    with oracle_ops():
        gen.eval()
        close(gen(z, c, noise_mode='const'), G['model/G_eval_const'], 1e-4, 2e-5)       # fused / grouped branch
        gen.train()
        close(gen(z, c, noise_mode='const'), G['model/G_train_const'], 1e-4, 2e-5)      # scale-activations branch
        torch.manual_seed(21)
        close(gen(z, c), G['model/G_train_random'], 1e-4, 2e-5)
        close(dis(real, c), G['model/D_real'], 1e-4, 2e-5)


@pytest.mark.parametrize('phase,which,gain', [('Gmain', 'G', 1), ('Greg', 'G', 4), ('Dmain', 'D', 1), ('Dreg', 'D', 16)])
def test_loss_phase_gradients(models, phase, which, gain):
    """Every parameter gradient of every phase, fixed seed, vs the reference's StyleGAN2Loss on the same weights."""
    G, gen, dis = models
    z, c, real = t(G['model/z']), t(G['model/c']), t(G['model/real'])
    aug = augment.AugmentPipe(**gg.AUG_KW).train().requires_grad_(False)
    aug.p.copy_(torch.as_tensor(0.6))
    loss = loss_mod.StyleGAN2Loss(device=torch.device('cpu'), G=gen, D=dis, augment_pipe=aug, **gg.LOSS_KW)
    # pl_mean evolves across the reference's phase sequence: Gmain (0.37) -> Greg updates it
    loss.pl_mean.copy_(torch.as_tensor(0.37))
    module = gen if which == 'G' else dis
    module.requires_grad_(True)
    for p in module.parameters():
        p.grad = None
    with oracle_ops():
        torch.manual_seed(100)
        loss.accumulate_gradients(phase=phase, real_img=real, real_c=c, gen_z=z, gen_c=c, gain=gain, cur_nimg=0)
    module.requires_grad_(False)
    checked = 0
    for name, p in module.named_parameters():
        key = f'loss/{phase}/{name}'
        if key in G:
            ref = G[key]
            if p.grad is None:
                # The reference's CPU `ref` path leaves exact zeros where its CUDA path (and ours) returns no gradient:
                # lrelu has no second derivative, so bias gradients of the R1 pass vanish (OPS/bias_act.py:186-203).
                assert not np.any(ref), name
                continue
            scale = max(np.abs(ref).max(), 1e-6)
            np.testing.assert_allclose(p.grad.numpy(), ref, rtol=2e-3, atol=2e-4 * scale, err_msg=name)
            checked += 1
        else:
            assert p.grad is None, name
    assert checked > 10
    if phase == 'Greg':
        close(loss.pl_mean, G['loss/Greg/pl_mean'], 1e-4, 1e-6)
