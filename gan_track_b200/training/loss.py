"""StyleGAN2 loss: non-saturating logistic loss with lazy R1 and path-length regularisation.

Same class, constructor and `accumulate_gradients(phase, real_img, real_c, gen_z, gen_c, gain, cur_nimg)` contract as
the reference (S3/training/loss.py:22-139); this file decides which forward / backward / double-backward graphs
exist, the op layer below decides how they run:
    Gmain  G -> [blur] -> augment -> D, softplus(-logits), backward into G
    Greg   path length: d(img . noise)/d(ws) with create_graph, penalty (|J| - pl_mean)^2, backward (2nd order through G)
    Dmain  fake: G (no grad) -> D, softplus(logits);  real: D, softplus(-logits)
    Dreg   R1: d(sum logits)/d(real_img) with create_graph, gamma/2 |.|^2, backward (2nd order through D and the ADA pipe)
RNG draws (style-mixing cutoff, second mapping pass, path-length noise) are issued in the reference's order so that a
fixed seed reproduces the reference's sequence.
"""
import numpy as np
import torch

from ..torch_utils import training_stats
from ..torch_utils.ops import conv2d_gradfix, rgb, upfirdn2d


class Loss:
    def accumulate_gradients(self, phase, real_img, real_c, gen_z, gen_c, gain, cur_nimg):
        raise NotImplementedError()


class StyleGAN2Loss(Loss):
    def __init__(self, device, G, D, augment_pipe=None, r1_gamma=10, style_mixing_prob=0, pl_weight=0, pl_batch_shrink=2,
                 pl_decay=0.01, pl_no_weight_grad=False, blur_init_sigma=0, blur_fade_kimg=0, allow_aug_debug_print=False,
                 merge_d_passes=False, merge_mapping_passes=False):
        super().__init__()
        self.device = device
        self.G = G
        self.D = D
        self.augment_pipe = augment_pipe
        self.r1_gamma = r1_gamma
        self.style_mixing_prob = style_mixing_prob
        self.pl_weight = pl_weight
        self.pl_batch_shrink = pl_batch_shrink
        self.pl_decay = pl_decay
        self.pl_no_weight_grad = pl_no_weight_grad
        self.pl_mean = torch.zeros([], device=device)
        self.blur_init_sigma = blur_init_sigma
        self.blur_fade_kimg = blur_fade_kimg
        self.allow_aug_debug_print = allow_aug_debug_print   # accepted for interface parity; debug plotting is not part of the path
        # Dmain scores the generated and the real batch with the same D; with merge_d_passes the two forward / backward
        # passes of the reference (S3/training/loss.py:102-139) run as one pass over the concatenated batch (same sums, see
        # _d_main_merged); off by default = the reference's schedule
        self.merge_d_passes = merge_d_passes
        # style mixing maps a second latent batch (S3/training/loss.py:46-50): with merge_mapping_passes both batches go through
        # the mapping network in one pass (rows are independent; w_avg tracks the first batch only)
        self.merge_mapping_passes = merge_mapping_passes

    def run_G(self, z, c, update_emas=False):
        if self.style_mixing_prob > 0 and self.merge_mapping_passes:
            # same draws in the same order as below (the mapping network itself draws nothing); z and the mixing latents go
            # through the mapping network as one batch, w_avg tracks the first half only
            num_ws = self.G.mapping.num_ws
            cutoff = torch.empty([], dtype=torch.int64, device=z.device).random_(1, num_ws)
            cutoff = torch.where(torch.rand([], device=z.device) < self.style_mixing_prob, cutoff, torch.full_like(cutoff, num_ws))
            both = self.G.mapping(torch.cat([z, torch.randn_like(z)]), torch.cat([c, c]), update_emas=update_emas, ema_rows=z.shape[0])
            ws, ws2 = both[:z.shape[0]], both[z.shape[0]:]
            layer = torch.arange(num_ws, device=z.device).reshape(1, -1, 1)
            ws = torch.where(layer < cutoff, ws, ws2)
            return self.G.synthesis(ws, update_emas=update_emas), ws
        ws = self.G.mapping(z, c, update_emas=update_emas)
        if self.style_mixing_prob > 0:
            cutoff = torch.empty([], dtype=torch.int64, device=ws.device).random_(1, ws.shape[1])
            cutoff = torch.where(torch.rand([], device=ws.device) < self.style_mixing_prob, cutoff, torch.full_like(cutoff, ws.shape[1]))
            # ws[:, cutoff:] = mapping(z2)[:, cutoff:] of the reference (S3/training/loss.py:48), written as a select so that
            # the random cutoff never leaves the device (slicing with a tensor index is a host sync, which would also make the
            # phase impossible to capture in a CUDA graph); same random draws in the same order
            ws2 = self.G.mapping(torch.randn_like(z), c, update_emas=False)
            layer = torch.arange(ws.shape[1], device=ws.device).reshape(1, -1, 1)
            ws = torch.where(layer < cutoff, ws, ws2)
        img = self.G.synthesis(ws, update_emas=update_emas)
        return img, ws

    def _pre_D(self, img, blur_sigma=0, allow_aug_debug_print=False):
        blur_size = np.floor(blur_sigma * 3)
        if blur_size > 0:
            f = torch.arange(-blur_size, blur_size + 1, device=img.device).div(blur_sigma).square().neg().exp2()
            img = upfirdn2d.filter2d(img, f / f.sum())
        if self.augment_pipe is not None:
            img = self.augment_pipe(img, allow_aug_debug_print)
        return img

    def run_D(self, img, c, blur_sigma=0, update_emas=False, allow_aug_debug_print=False):
        return self.D(self._pre_D(img, blur_sigma, allow_aug_debug_print), c, update_emas=update_emas)

    # -- the four sub-graphs ------------------------------------------------------------------------------------

    def _g_main(self, gen_z, gen_c, gain, blur_sigma):
        """Generator wants D to score its images high: softplus(-D(G(z)))."""
        fake, _ = self.run_G(gen_z, gen_c)
        logits = self.run_D(fake, gen_c, blur_sigma=blur_sigma)
        training_stats.report('Loss/scores/fake', logits)
        training_stats.report('Loss/signs/fake', logits.sign())
        loss = torch.nn.functional.softplus(-logits)
        training_stats.report('Loss/G/loss', loss)
        loss.mean().mul(gain).backward()

    def _g_pathlen(self, gen_z, gen_c, gain):
        """Path-length regulariser on a shrunken batch: second-order backward through the synthesis network."""
        n = gen_z.shape[0] // self.pl_batch_shrink
        with rgb.op_by_op_torgb():          # this pass differentiates G's backward: keep ToRGB in its differentiable op-by-op form
            fake, ws = self.run_G(gen_z[:n], gen_c[:n])
        probe = torch.randn_like(fake) / np.sqrt(fake.shape[2] * fake.shape[3])
        with conv2d_gradfix.no_weight_gradients(self.pl_no_weight_grad):
            jac, = torch.autograd.grad(outputs=[(fake * probe).sum()], inputs=[ws], create_graph=True, only_inputs=True)
        lengths = jac.square().sum(2).mean(1).sqrt()
        mean = self.pl_mean.lerp(lengths.mean(), self.pl_decay)
        self.pl_mean.copy_(mean.detach())
        penalty = (lengths - mean).square()
        training_stats.report('Loss/pl_penalty', penalty)
        loss = penalty * self.pl_weight
        training_stats.report('Loss/G/reg', loss)
        loss.mean().mul(gain).backward()

    def _d_fake(self, gen_z, gen_c, gain, blur_sigma):
        """Discriminator wants generated images scored low: softplus(D(G(z))).  Returns the per-sample loss for stats."""
        fake, _ = self.run_G(gen_z, gen_c, update_emas=True)
        logits = self.run_D(fake, gen_c, blur_sigma=blur_sigma, update_emas=True)
        training_stats.report('Loss/scores/fake', logits)
        training_stats.report('Loss/signs/fake', logits.sign())
        loss = torch.nn.functional.softplus(logits)
        loss.mean().mul(gain).backward()
        return loss

    def _d_real(self, real_img, real_c, gain, blur_sigma, loss_fake, with_main, with_r1):
        """Real images scored high (with_main) and / or the lazy R1 gradient penalty (with_r1)."""
        real = real_img.detach().requires_grad_(with_r1)
        logits = self.run_D(real, real_c, blur_sigma=blur_sigma, allow_aug_debug_print=self.allow_aug_debug_print)
        training_stats.report('Loss/scores/real', logits)
        training_stats.report('Loss/signs/real', logits.sign())
        total = 0
        if with_main:
            loss_real = torch.nn.functional.softplus(-logits)
            training_stats.report('Loss/D/loss', loss_fake + loss_real)
            total = total + loss_real
        if with_r1:
            with conv2d_gradfix.no_weight_gradients():
                g, = torch.autograd.grad(outputs=[logits.sum()], inputs=[real], create_graph=True, only_inputs=True)
            penalty = g.square().sum([1, 2, 3])
            loss_r1 = penalty * (self.r1_gamma / 2)
            training_stats.report('Loss/r1_penalty', penalty)
            training_stats.report('Loss/D/reg', loss_r1)
            total = total + loss_r1
        total.mean().mul(gain).backward()

    def _mbstd_group(self, n):
        """Group size the discriminator's MinibatchStdLayer uses for a batch of n, None if it has no such layer."""
        layer = getattr(getattr(self.D, 'b4', None), 'mbstd', None)
        if layer is None:
            return 1
        if layer.group_size is None:
            return None             # one group = the whole batch: merging would change the statistics
        return min(int(layer.group_size), int(n))

    def _can_merge_d(self, real_img, gen_z):
        if not self.merge_d_passes or real_img.shape[0] != gen_z.shape[0]:
            return False
        g = self._mbstd_group(gen_z.shape[0])
        return g is not None and gen_z.shape[0] % g == 0

    def _d_main_merged(self, real_img, real_c, gen_z, gen_c, gain, blur_sigma):
        """Dmain as ONE discriminator pass over [generated, real]: loss = mean softplus(D(fake)) + mean softplus(-D(real)),
        one backward.  Per-sample work in D is independent of the batch except MinibatchStdLayer, whose group j is the samples
        {j, j + n, j + 2n, ...} (n = batch / G, networks_stylegan2.py:650-665); the two batches are interleaved in chunks of
        batch/G so that every group of the merged batch is exactly a group of one of the separate passes.  Random draws
        (G noise, ADA parameters for fake, then for real) happen in the reference's order; parameter gradients are the same
        sums, accumulated in one pass instead of two."""
        n = gen_z.shape[0]
        g = self._mbstd_group(n)
        fake, _ = self.run_G(gen_z, gen_c, update_emas=True)
        fake = self._pre_D(fake, blur_sigma)
        real = self._pre_D(real_img.detach(), blur_sigma, self.allow_aug_debug_print)

        def interleave(a, b):
            return torch.cat([t for pair in zip(a.chunk(g), b.chunk(g)) for t in pair])

        logits = self.D(interleave(fake, real), interleave(gen_c, real_c), update_emas=True)
        parts = logits.chunk(2 * g)
        logits_fake, logits_real = torch.cat(parts[0::2]), torch.cat(parts[1::2])
        training_stats.report('Loss/scores/fake', logits_fake)
        training_stats.report('Loss/signs/fake', logits_fake.sign())
        loss_fake = torch.nn.functional.softplus(logits_fake)
        training_stats.report('Loss/scores/real', logits_real)
        training_stats.report('Loss/signs/real', logits_real.sign())
        loss_real = torch.nn.functional.softplus(-logits_real)
        training_stats.report('Loss/D/loss', loss_fake + loss_real)
        (loss_fake.mean() + loss_real.mean()).mul(gain).backward()

    def accumulate_gradients(self, phase, real_img, real_c, gen_z, gen_c, gain, cur_nimg):
        assert phase in ['Gmain', 'Greg', 'Gboth', 'Dmain', 'Dreg', 'Dboth']
        if self.pl_weight == 0:
            phase = {'Greg': 'none', 'Gboth': 'Gmain'}.get(phase, phase)
        if self.r1_gamma == 0:
            phase = {'Dreg': 'none', 'Dboth': 'Dmain'}.get(phase, phase)
        blur_sigma = 0
        if self.blur_fade_kimg > 0:
            blur_sigma = max(1 - cur_nimg / (self.blur_fade_kimg * 1e3), 0) * self.blur_init_sigma
        if phase in ('Gmain', 'Gboth'):
            self._g_main(gen_z, gen_c, gain, blur_sigma)
        if phase in ('Greg', 'Gboth'):
            self._g_pathlen(gen_z, gen_c, gain)
        loss_fake = 0
        if phase == 'Dmain' and self._can_merge_d(real_img, gen_z):
            self._d_main_merged(real_img, real_c, gen_z, gen_c, gain, blur_sigma)
            return
        if phase in ('Dmain', 'Dboth'):
            loss_fake = self._d_fake(gen_z, gen_c, gain, blur_sigma)
        if phase in ('Dmain', 'Dreg', 'Dboth'):
            self._d_real(real_img, real_c, gain, blur_sigma, loss_fake, with_main=phase in ('Dmain', 'Dboth'), with_r1=phase in ('Dreg', 'Dboth'))
