"""CPU: the training step (phases, lazy-reg Adam, gradient exchange, EMA, ADA) with the oracle's primitive ops;
world_size-2 `gloo` run for the data-parallel path (bucketed / overlapped reducer vs the plain flat all-reduce)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT
from oracle.backend import oracle_ops
from gan_track_b200.training import training_loop as tl


def tiny_cfg(num_gpus=1, batch=4):
    cfg = tl.claro_config(resolution=16, batch=batch, num_gpus=num_gpus, cbase=128, cmax=8, map_depth=2)
    cfg.G_kwargs.z_dim = cfg.G_kwargs.w_dim = 16
    return cfg


def batch_for(rank, n):
    g = torch.Generator().manual_seed(7 + rank)
    real = torch.rand([n, 1, 16, 16], generator=g) * 255
    c = torch.nn.functional.one_hot(torch.randint(0, 2, [n], generator=g), 2).float()
    return real, c


def test_single_process_step_schedule_and_updates():
    with oracle_ops():
        tr = tl.Trainer(tiny_cfg(), rank=0, device='cpu', overlap=False)
        g0 = [p.clone() for p in tr.G.parameters()]
        d0 = [p.clone() for p in tr.D.parameters()]
        e0 = [p.clone() for p in tr.G_ema.parameters()]
        real, c = batch_for(0, 4)
        for _ in range(5):
            tr.train_step(real, c)
    assert tr.phase_counts == {'Gmain': 5, 'Greg': 2, 'Dmain': 5, 'Dreg': 1}          # intervals 1 / 4 / 1 / 16
    assert tr.cur_nimg == 20 and tr.batch_idx == 5
    assert any((a != b).any() for a, b in zip(g0, tr.G.parameters()))
    assert any((a != b).any() for a, b in zip(d0, tr.D.parameters()))
    assert any((a != b).any() for a, b in zip(e0, tr.G_ema.parameters()))
    for p in list(tr.G.parameters()) + list(tr.D.parameters()):
        assert torch.isfinite(p).all()
    # lazy regularisation: lr * c and beta ** c with c = interval / (interval + 1)   (training_loop_mi_multimodal.py:248-255)
    gopt = tr.phases[0].opt.param_groups[0]
    assert np.isclose(gopt['lr'], 0.0025 * 4 / 5) and np.isclose(gopt['betas'][1], 0.99 ** (4 / 5))
    dopt = tr.phases[2].opt.param_groups[0]
    assert np.isclose(dopt['lr'], 0.0025 * 16 / 17) and np.isclose(dopt['betas'][1], 0.99 ** (16 / 17))
    assert tr.phases[0].opt is tr.phases[1].opt and tr.phases[2].opt is tr.phases[3].opt
    assert float(tr.augment_pipe.p) >= 0


def _dmain_grads(merge, batch=8, aug_p=0.7):
    from gan_track_b200.torch_utils import training_stats
    with oracle_ops():
        tr = tl.Trainer(tiny_cfg(batch=batch), rank=0, device='cpu', overlap=False, merge_d_passes=merge)
        tr.augment_pipe.p.fill_(aug_p)
        real, c = batch_for(0, batch)
        g = torch.Generator().manual_seed(11)
        z = torch.randn([batch, 16], generator=g)
        gc = torch.nn.functional.one_hot(torch.randint(0, 2, [batch], generator=g), 2).float()
        tr.D.requires_grad_(True)
        tr.G.requires_grad_(False)
        torch.manual_seed(5)
        table = training_stats._table(torch.device('cpu'))
        table.zero_()
        tr.loss.accumulate_gradients(phase='Dmain', real_img=real / 127.5 - 1, real_c=c, gen_z=z, gen_c=gc, gain=1, cur_nimg=0)
        grads = {n: p.grad.clone() for n, p in tr.D.named_parameters() if p.grad is not None}
        stats = {k: table[row].clone() for k, row in training_stats._name_to_row.items() if k.startswith('Loss/')}
    return grads, stats


def test_dmain_merged_pass_equals_two_passes():
    """One discriminator pass over the interleaved [generated, real] batch (MinibatchStd groups preserved, same random
    draws) gives the parameter gradients and the logged statistics of the reference's two passes."""
    ga, sa = _dmain_grads(False)
    gb, sb = _dmain_grads(True)
    assert ga.keys() == gb.keys() and len(ga) > 10
    for n in ga:
        scale = float(ga[n].abs().max().clamp_min(1e-12))
        assert float((ga[n] - gb[n]).abs().max()) <= 1e-5 * scale + 1e-9, n
    assert sa.keys() == sb.keys() and {'Loss/scores/fake', 'Loss/scores/real', 'Loss/D/loss'} <= set(sa.keys())
    for k in sa:
        assert torch.allclose(sa[k], sb[k], rtol=1e-5, atol=1e-6), k
    # batches that do not split into whole MinibatchStd groups keep the two-pass schedule
    with oracle_ops():
        tr = tl.Trainer(tiny_cfg(batch=4), rank=0, device='cpu', overlap=False, merge_d_passes=True)
        assert tr.loss._can_merge_d(torch.zeros(4, 1, 16, 16), torch.zeros(4, 16))
        assert not tr.loss._can_merge_d(torch.zeros(4, 1, 16, 16), torch.zeros(2, 16))
        tr.D.b4.mbstd.group_size = None
        assert not tr.loss._can_merge_d(torch.zeros(4, 1, 16, 16), torch.zeros(4, 16))


def test_style_mixing_single_mapping_pass_equals_two_passes():
    """run_G with both latent batches through the mapping network at once: same images, same ws, same w_avg update."""
    out = []
    for merge in (False, True):
        with oracle_ops():
            tr = tl.Trainer(tiny_cfg(batch=4), rank=0, device='cpu', overlap=False)
            tr.loss.merge_mapping_passes = merge
            g = torch.Generator().manual_seed(13)
            z = torch.randn([4, 16], generator=g)
            c = torch.nn.functional.one_hot(torch.randint(0, 2, [4], generator=g), 2).float()
            torch.manual_seed(21)
            img, ws = tr.loss.run_G(z, c, update_emas=True)
            out.append((img.detach(), ws.detach(), tr.G.mapping.w_avg.clone()))
    for a, b, name in zip(out[0], out[1], ['img', 'ws', 'w_avg']):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6), name
    assert out[0][2].abs().sum() > 0


def _worker(rank, world, port, overlap, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.distributed.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    from gan_track_b200.torch_utils import training_stats
    training_stats.init_multiprocessing(rank=rank, sync_device=torch.device('cpu'))
    with oracle_ops():
        tr = tl.Trainer(tiny_cfg(num_gpus=world, batch=8), rank=rank, device='cpu', overlap=overlap)
        real, c = batch_for(rank, 4)
        for _ in range(5):
            tr.train_step(real, c)
        tr.check_consistency()                       # replicas bit-identical (misc.check_ddp_consistency)
    if overlap:
        assert tr.phases[0].grad_set, 'bucketed reducer never learned the gradient set'
    torch.save({'G': tr.G.state_dict(), 'D': tr.D.state_dict(), 'p': tr.augment_pipe.p}, os.path.join(out_dir, f'r{rank}_{int(overlap)}.pt'))
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize('overlap', [False, True])
def test_two_rank_gloo_replicas_stay_identical(tmp_path, overlap):
    port = 29500 + (os.getpid() % 500) + (7 if overlap else 0)
    mp.spawn(_worker, args=(2, port, overlap, str(tmp_path)), nprocs=2, join=True)
    a = torch.load(os.path.join(tmp_path, f'r0_{int(overlap)}.pt'))
    b = torch.load(os.path.join(tmp_path, f'r1_{int(overlap)}.pt'))
    for k in a['G']:
        if not k.endswith('w_avg'):
            assert torch.equal(a['G'][k], b['G'][k]), k
    for k in a['D']:
        assert torch.equal(a['D'][k], b['D'][k]), k
    assert torch.equal(a['p'], b['p'])


def test_bucketed_reducer_matches_flat_allreduce(tmp_path):
    """Same seeds, overlap on vs off: identical parameters after 5 steps (the bucketing only reorders communication)."""
    for overlap in (False, True):
        port = 30100 + (os.getpid() % 500) + (11 if overlap else 0)
        mp.spawn(_worker, args=(2, port, overlap, str(tmp_path)), nprocs=2, join=True)
    a = torch.load(os.path.join(tmp_path, 'r0_0.pt'))
    b = torch.load(os.path.join(tmp_path, 'r0_1.pt'))
    for k in a['G']:
        assert torch.equal(a['G'][k], b['G'][k]), k
    for k in a['D']:
        assert torch.equal(a['D'][k], b['D'][k]), k
