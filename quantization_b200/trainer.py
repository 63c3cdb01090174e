"""`QuantizerTrainer`: the reference's two-phase training driver (quantization/quantization.py:577-742) on top of the
B200-native `Quantizer`.  Same constructor, `step` / `done` / `get_quantizer`, same RNG draws (one
`random.random()` per step, :651), same optimiser / scheduler settings (:722-730), same log lines (:656-675).
The index refinement and decode inside `compute_loss` run in libmcq.so; the loss arithmetic, backward, Adam and
StepLR are stock PyTorch, as in the reference."""
import logging
import random
import time

import torch

from .quantizer import Quantizer


class QuantizerTrainer(object):
    def __init__(self, dim: int, bytes_per_frame: int, device: torch.device, phase_one_iters: int = 10000,
                 phase_two_iters: int = 10000, lr: float = 0.005):
        super().__init__()
        assert bytes_per_frame in [1, 2, 4, 8, 16, 32]
        self.phase_one_iters = phase_one_iters
        self.phase_two_iters = phase_two_iters
        self.cur_iter = 0
        self.lr = lr
        self.two_iter_prob = 0.5
        # phase 1: codebook_size 16 with twice the codebooks; phase 2: pairs multiplied out to codebook_size 256
        self.quantizer = Quantizer(dim=dim, codebook_size=16, num_codebooks=bytes_per_frame * 2).to(device)
        self.start_time = time.time()
        self._init_optimizer()

    def done(self) -> bool:
        ans = self.cur_iter > self.phase_one_iters + self.phase_two_iters
        if ans:
            elapsed_time = time.time() - self.start_time
            logging.info(f"Elapsed time, training model of dim={self.quantizer.dim}, "
                         f"num_codebooks={self.quantizer.num_codebooks}, "
                         f"codebook_size={self.quantizer.codebook_size}, is: {elapsed_time:.2f} seconds.")
        return ans

    def step(self, x: torch.Tensor) -> None:
        x = x.reshape(-1, self.quantizer.dim)
        num_iters = 2 if random.random() < self.two_iter_prob else 1
        (reconstruction_loss, logprob_loss, logits_entropy_loss,
         index_entropy_loss) = self.quantizer.compute_loss(x, num_iters)

        if self.cur_iter % 200 == 0:
            det_losses = [float('%.3f' % self.quantizer.compute_loss(x, j)[0].item()) for j in range(6)]
            phase = 1 if self.cur_iter <= self.phase_one_iters else 2
            i = self.cur_iter - self.phase_one_iters if phase > 1 else self.cur_iter
            logging.info(f"phase={phase}/2, iter={i}, "
                         f"dim,nc,csz={self.quantizer.dim},{self.quantizer.num_codebooks},"
                         f"{self.quantizer.codebook_size}, "
                         f"loss_per_iter={det_losses}, "
                         f"logprob_loss={logprob_loss.item():.3f}, "
                         f"logits_entropy_loss={logits_entropy_loss.item():.3f}, "
                         f"index_entropy_loss={index_entropy_loss.item():.3f}")

        if self.cur_iter % 2000 == 0 and self.cur_iter > 0:
            correlations = self.quantizer.compute_codebook_correlations()
            logging.info(f"correlations = {correlations}")

        entropy_scale = 0.01
        tot_loss = reconstruction_loss + logprob_loss + logits_entropy_loss * entropy_scale
        tot_loss.backward()
        self.optim.step()
        self.optim.zero_grad()
        self.scheduler.step()

        if self.cur_iter == self.phase_one_iters:
            self._begin_second_phase()
        self.cur_iter += 1

    def _init_optimizer(self):
        self.optim = torch.optim.Adam(self.quantizer.parameters(), lr=self.lr, betas=(0.9, 0.98), eps=1e-9,
                                      weight_decay=1.0e-06)
        step_size = (self.phase_one_iters if self.cur_iter == 0 else self.phase_two_iters) / 4
        self.scheduler = torch.optim.lr_scheduler.StepLR(self.optim, step_size=step_size, gamma=0.5)

    def _begin_second_phase(self):
        self.quantizer = self.quantizer.get_product_quantizer()
        self.lr *= 0.5
        self._init_optimizer()

    def get_quantizer(self) -> Quantizer:
        assert self.cur_iter >= self.phase_one_iters + self.phase_two_iters
        return self.quantizer
