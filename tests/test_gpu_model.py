"""GPU (-m gpu): whole networks and loss phases on the CUDA path vs golden vectors from the real reference (fp32) and
fp16-vs-fp32 agreement at the north star's 1e-2."""
import numpy as np
import pytest
import torch

from oracle import gen_golden as gg

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from gan_track_b200.training import augment, loss as loss_mod, networks_stylegan2 as nets

DEV = 'cuda'


def t(a):
    return torch.from_numpy(np.asarray(a)).to(DEV)


def rel_err(a, b):
    a, b = a.detach().double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def _load(module, G, prefix):
    module.load_state_dict({k[len(prefix):]: torch.from_numpy(G[k]) for k in G.keys(prefix)}, strict=True)
    return module


@pytest.fixture(scope='module')
def models32(golden):
    G = golden('model.npz')
    kwg = dict(gg.G_KW, num_fp16_res=0, conv_clamp=None)
    kwd = dict(gg.D_KW, num_fp16_res=0, conv_clamp=None)
    # the golden models were built with the default fp16 settings but evaluated on the CPU, where the reference forces
    # fp32 (networks_stylegan2.py:419-420, 607-608); conv_clamp=256 never binds at these magnitudes but keep it exact:
    kwg['conv_clamp'] = 256
    kwd['conv_clamp'] = 256
    gen = _load(nets.Generator(**kwg), G, 'model/G/').train().requires_grad_(False).to(DEV)
    dis = _load(nets.Discriminator(**kwd), G, 'model/D/').train().requires_grad_(False).to(DEV)
    return G, gen, dis


def test_networks_fp32_vs_reference_golden(models32):
    G, gen, dis = models32
    z, c, real = t(G['model/z']), t(G['model/c']), t(G['model/real'])
    gen.eval()
    # the reference's eval-mode output (its grouped fused_modconv branch) vs both of our evaluations of that branch: the
    # default shared-weight route and the grouped convolution
    assert rel_err(gen(z, c, noise_mode='const'), G['model/G_eval_const']) <= 2e-5
    nets.grouped_fused_modconv = True
    try:
        assert rel_err(gen(z, c, noise_mode='const'), G['model/G_eval_const']) <= 2e-5
        assert rel_err(gen(z[:1], c[:1], noise_mode='const'), G['model/G_eval_const'][:1]) <= 2e-5
    finally:
        nets.grouped_fused_modconv = False
    assert rel_err(gen(z[:1], c[:1], noise_mode='const'), G['model/G_eval_const'][:1]) <= 2e-5       # batch 1
    gen.train()
    assert rel_err(gen(z, c, noise_mode='const'), G['model/G_train_const']) <= 2e-5
    assert rel_err(dis(real, c), G['model/D_real']) <= 2e-5


def test_networks_fp16_vs_fp32(golden, models32):
    G, gen32, dis32 = models32
    gen16 = _load(nets.Generator(**gg.G_KW), G, 'model/G/').train().requires_grad_(False).to(DEV)
    dis16 = _load(nets.Discriminator(**gg.D_KW), G, 'model/D/').train().requires_grad_(False).to(DEV)
    z, c, real = t(G['model/z']), t(G['model/c']), t(G['model/real'])
    assert any(b.use_fp16 for b in gen16.synthesis.children())
    assert rel_err(gen16(z, c, noise_mode='const'), gen32(z, c, noise_mode='const')) <= 1e-2
    assert rel_err(dis16(real, c), dis32(real, c)) <= 1e-2


@pytest.mark.parametrize('phase,which,gain', [('Gmain', 'G', 1), ('Greg', 'G', 4), ('Dmain', 'D', 1), ('Dreg', 'D', 16)])
def test_loss_phase_gradients_fp32(models32, phase, which, gain):
    """CUDA RNG differs from the CPU generator the golden run used, so randomness is removed: ADA in known-answer mode is
    not available through the loss, hence the pipe is disabled and noise strengths are what the golden model has; style
    mixing / pl noise make Gmain / Greg seed-dependent, so those two are compared CUDA-vs-oracle on identical draws by
    running the oracle backend on the CPU copy with draws replayed from the device."""
    G, gen, dis = models32
    z, c, real = t(G['model/z']), t(G['model/c']), t(G['model/real'])
    loss = loss_mod.StyleGAN2Loss(device=torch.device(DEV), G=gen, D=dis, augment_pipe=None, r1_gamma=0.4096, style_mixing_prob=0, pl_weight=2,
                                  pl_no_weight_grad=True)
    loss.pl_mean.copy_(torch.as_tensor(0.37))
    module = gen if which == 'G' else dis
    # deterministic path: const noise for G (monkeypatch forward kwargs), fixed pl noise via manual seeding of CUDA + replay on CPU
    import copy
    from oracle.backend import oracle_ops
    gen_cpu, dis_cpu = copy.deepcopy(gen).cpu(), copy.deepcopy(dis).cpu()
    loss_cpu = loss_mod.StyleGAN2Loss(device=torch.device('cpu'), G=gen_cpu, D=dis_cpu, augment_pipe=None, r1_gamma=0.4096, style_mixing_prob=0,
                                      pl_weight=2, pl_no_weight_grad=True)
    loss_cpu.pl_mean.copy_(torch.as_tensor(0.37))

    # replay device randomness on the CPU: patch torch.randn / randn_like inside the CPU run to pop recorded draws
    draws = []
    orig_randn, orig_randn_like = torch.randn, torch.randn_like

    def rec_randn(*a, **k):
        r = orig_randn(*a, **k)
        draws.append(r.detach().cpu())
        return r

    def rec_randn_like(x, **k):
        r = orig_randn_like(x, **k)
        draws.append(r.detach().cpu())
        return r

    module.requires_grad_(True)
    for p in module.parameters():
        p.grad = None
    torch.randn, torch.randn_like = rec_randn, rec_randn_like
    try:
        loss.accumulate_gradients(phase=phase, real_img=real, real_c=c, gen_z=z, gen_c=c, gain=gain, cur_nimg=0)
    finally:
        torch.randn, torch.randn_like = orig_randn, orig_randn_like
    module.requires_grad_(False)

    it = iter(draws)
    module_cpu = gen_cpu if which == 'G' else dis_cpu
    module_cpu.requires_grad_(True)
    torch.randn = lambda *a, **k: next(it).clone()
    torch.randn_like = lambda x, **k: next(it).clone()
    try:
        with oracle_ops():
            loss_cpu.accumulate_gradients(phase=phase, real_img=real.cpu(), real_c=c.cpu(), gen_z=z.cpu(), gen_c=c.cpu(), gain=gain, cur_nimg=0)
    finally:
        torch.randn, torch.randn_like = orig_randn, orig_randn_like
    module_cpu.requires_grad_(False)

    checked = 0
    for (name, p), (_, q) in zip(module.named_parameters(), module_cpu.named_parameters()):
        if q.grad is None or not bool(q.grad.abs().max() > 0):
            continue
        assert p.grad is not None, name
        scale = float(q.grad.abs().max())
        err = float((p.grad.cpu() - q.grad).abs().max()) / scale
        assert err <= 5e-4, f'{phase} {name}: {err:.2e}'
        checked += 1
    assert checked > 10


def test_augment_pipe_cuda_known_answers(golden):
    G = golden('augment.npz')
    x = t(G['augment/img'])
    pipe = augment.AugmentPipe(**gg.AUG_KW).to(DEV)
    pipe.p.copy_(torch.as_tensor(0.7))
    for pct in (0.1, 0.5, 0.9):
        out = pipe(x, False, debug_percentile=pct)
        assert rel_err(out, G[f'augment/claro/pct{pct}']) <= 5e-5, pct
    pipe = augment.AugmentPipe(**{k: v for k, v in gg.AUG_KW_FULL.items() if k not in ('noise', 'cutout')}).to(DEV)
    pipe.p.copy_(torch.as_tensor(0.7))
    ref_pipe_keys = [0.1, 0.5, 0.9]
    for pct in ref_pipe_keys:
        out = pipe(x, False, debug_percentile=pct)
        assert torch.isfinite(out).all()


def test_trainer_cuda_graph_mode_runs_and_captures():
    """Graph mode: every phase is captured at its second occurrence and replayed; parameters stay finite and move."""
    from gan_track_b200.training import training_loop as tl
    cfg = tl.claro_config(resolution=32, batch=4, num_gpus=1, cbase=1024, cmax=64, map_depth=2)
    trainer = tl.Trainer(cfg, rank=0, device='cuda', use_graphs=True)
    real = torch.rand(4, 1, 32, 32) * 255
    c = torch.nn.functional.one_hot(torch.tensor([0, 1, 1, 0]), 2).float()
    before = [p.detach().clone() for p in trainer.G.parameters()]
    for _ in range(18):
        trainer.train_step(real, c)
    torch.cuda.synchronize()
    assert all(ph.graphs is not None for ph in trainer.phases), 'Gmain/Greg/Dmain/Dreg must all be captured after 17 iterations'
    assert trainer.phase_counts == {'Gmain': 18, 'Greg': 5, 'Dmain': 18, 'Dreg': 2}
    moved = 0
    for p, b in zip(trainer.G.parameters(), before):
        assert torch.isfinite(p).all()
        moved += int(not torch.equal(p, b))
    assert moved > 0
    for p in trainer.D.parameters():
        assert torch.isfinite(p).all()


def test_full_width_training_step_needs_no_library_convolution():
    """The 256x256 configuration of BASELINE config 2 (cbase 16384: 64 ... 512 channels, fp16 top-4 resolutions, ADA) with
    `conv_backend.allow_library = False`: every convolution of all four phases -- fp16 blocks, the true-fp32 4^2 - 16^2 blocks (fp16 x 3),
    forward, data and weight gradients, both double backwards -- must be taken by this package's tcgen05 kernels, else the step raises."""
    from gan_track_b200.torch_utils.ops import conv_backend
    from gan_track_b200.training import training_loop as tl
    cfg = tl.claro_config(resolution=256, batch=4, num_gpus=1)
    trainer = tl.Trainer(cfg, rank=0, device='cuda', use_graphs=False)
    real = torch.rand(4, 1, 256, 256) * 255
    c = torch.nn.functional.one_hot(torch.tensor([0, 1, 1, 0]), 2).float()
    before = dict(conv_backend.stats)
    old = conv_backend.allow_library
    conv_backend.allow_library = False
    try:
        trainer.train_step(real, c)              # iteration 0 runs Gmain, Greg, Dmain and Dreg
        torch.cuda.synchronize()
    finally:
        conv_backend.allow_library = old
    assert trainer.phase_counts == {'Gmain': 1, 'Greg': 1, 'Dmain': 1, 'Dreg': 1}
    assert conv_backend.stats['library'] == before['library'] and conv_backend.stats['library_wgrad'] == before['library_wgrad']
    assert conv_backend.stats['igemm'] > before['igemm'] + 100 and conv_backend.stats['igemm_wgrad'] > before['igemm_wgrad'] + 30
    for p in list(trainer.G.parameters()) + list(trainer.D.parameters()):
        assert torch.isfinite(p).all()


@pytest.mark.parametrize('fp32,tol', [(True, 2e-4), (False, 2e-2)])
def test_dmain_merged_pass_equals_two_passes_cuda(fp32, tol):
    """Dmain as one discriminator pass over the interleaved [generated, real] batch vs the reference's two passes, on the
    CUDA kernels (fp32 model and fp16 model on the tcgen05 path): same parameter gradients up to summation order.  The fp32
    bound is 2e-4, not 1e-5: weight gradients are sums over 10^5..10^6 mixed-sign terms whose grouping differs between one pass
    over 2N samples and two passes over N (observed up to 4e-5 on fromrgb.weight, varying with the library's algorithm choice);
    the exact equivalence of the schedule is pinned at 1e-5 by the CPU test (tests/test_training_step.py)."""
    from gan_track_b200.training import training_loop as tl
    grads = []
    for merge in (False, True):
        cfg = tl.claro_config(resolution=64, batch=16, num_gpus=1, cbase=4096, cmax=64, map_depth=2, fp32=fp32)
        tr = tl.Trainer(cfg, rank=0, device='cuda', use_graphs=False, merge_d_passes=merge)
        tr.augment_pipe.p.fill_(0.6)
        g = torch.Generator().manual_seed(3)
        real = (torch.rand([16, 1, 64, 64], generator=g) * 2 - 1).cuda()
        c = torch.nn.functional.one_hot(torch.randint(0, 2, [16], generator=g), 2).float().cuda()
        z = torch.randn([16, 512], generator=g).cuda()
        tr.D.requires_grad_(True)
        tr.G.requires_grad_(False)
        torch.manual_seed(9)
        tr.loss.accumulate_gradients(phase='Dmain', real_img=real, real_c=c, gen_z=z, gen_c=c.flip(0), gain=1, cur_nimg=0)
        grads.append({n: p.grad.detach().clone() for n, p in tr.D.named_parameters() if p.grad is not None})
    assert grads[0].keys() == grads[1].keys() and len(grads[0]) > 10
    for n in grads[0]:
        assert torch.isfinite(grads[1][n]).all(), n
        assert rel_err(grads[1][n], grads[0][n]) <= tol, (n, rel_err(grads[1][n], grads[0][n]))
