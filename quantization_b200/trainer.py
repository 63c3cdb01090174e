"""`QuantizerTrainer`: the reference's two-phase training driver (quantization/quantization.py:577-742) on top of the
B200-native `Quantizer`.  Same constructor, `step` / `done` / `get_quantizer`, same RNG draws (one
`random.random()` per step, :651), same optimiser / scheduler settings (:722-730), same log lines (:656-675).
The index refinement, the losses and their backward inside `compute_loss` run in libmcq.so; Adam and StepLR are stock
PyTorch, as in the reference.

A step is ~120 kernel launches of a few microseconds each, so on a B200 the HOST is the bound (2.2 ms of Python and
launch overhead around 1.5 ms of kernels at 65,536 frames).  The whole step -- compute_loss, backward, Adam -- is
therefore captured in a CUDA graph per (quantizer, batch shape, refinement passes, learning rate) after three eager
steps, and replayed; the diagnostics iterations, the phase switch and anything with a new shape run eagerly.
`MCQ_TRAINER_GRAPH=0` disables the capture."""
import logging
import os
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
        self._use_graph = torch.device(device).type == "cuda" and os.environ.get("MCQ_TRAINER_GRAPH", "1") != "0"
        self._graphs = {}   # key -> (CUDAGraph, static input, static losses, tensors the graph's kernels point into)
        self._warm = {}     # key -> eager steps seen
        self._init_optimizer()

    def done(self) -> bool:
        ans = self.cur_iter > self.phase_one_iters + self.phase_two_iters
        if ans:
            elapsed_time = time.time() - self.start_time
            logging.info(f"Elapsed time, training model of dim={self.quantizer.dim}, "
                         f"num_codebooks={self.quantizer.num_codebooks}, "
                         f"codebook_size={self.quantizer.codebook_size}, is: {elapsed_time:.2f} seconds.")
        return ans

    _GRAPH_WARMUP = 3   # eager steps before a configuration is captured (lazy initialisations happen there)
    _MAX_GRAPHS = 8

    def _graph_key(self, x: torch.Tensor, num_iters: int):
        return (id(self.quantizer), tuple(x.shape), x.dtype, num_iters, float(self.optim.param_groups[0]["lr"]))

    def _loss_and_update(self, x: torch.Tensor, num_iters: int):
        """compute_loss + backward + optimiser step (reference :653, :677-681); returns the four losses."""
        losses = self.quantizer.compute_loss(x, num_iters)
        entropy_scale = 0.01
        tot_loss = losses[0] + losses[1] + losses[2] * entropy_scale
        tot_loss.backward()
        self.optim.step()
        return losses

    def _eager_update(self, x: torch.Tensor, num_iters: int):
        """One eager step.  The gradients are cleared BEFORE the backward as well: after a capture / replay `p.grad`
        may still point at a graph's buffers holding the previous replay's gradients."""
        self.optim.zero_grad(set_to_none=True)
        losses = self._loss_and_update(x, num_iters)
        self.optim.zero_grad(set_to_none=True)
        return losses

    def _graphed_update(self, x: torch.Tensor, num_iters: int):
        key = self._graph_key(x, num_iters)
        if self._graphs and next(iter(self._graphs))[-1] != key[-1]:
            self._graphs.clear()  # the learning rate moved (StepLR): it is baked into the captured Adam kernels
            self._warm.clear()
        rec = self._graphs.get(key)
        if rec is None:
            seen = self._warm.get(key, 0)
            if seen < self._GRAPH_WARMUP or len(self._graphs) >= self._MAX_GRAPHS:
                if len(self._warm) > 256:  # ever-changing batch shapes: nothing worth capturing, keep the table small
                    self._warm.clear()
                self._warm[key] = seen + 1
                return self._eager_update(x, num_iters)
            q = self.quantizer
            static_x = x.clone()
            self.optim.zero_grad(set_to_none=True)
            torch.cuda.synchronize(x.device)
            graph = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(graph):
                    # detached: the caller only reads the values.  Keeping grad_fn alive would keep the captured
                    # autograd graph -- and its AccumulateGrad nodes, which are bound to the capture stream -- alive,
                    # and every later EAGER backward would accumulate on that stream instead of the current one
                    losses = tuple(l.detach() for l in self._loss_and_update(static_x, num_iters))
            except RuntimeError as e:  # not capturable in this environment: same kernels, launched eagerly from now on
                logging.warning(f"QuantizerTrainer: CUDA-graph capture failed ({e}); continuing without graphs")
                self._use_graph = False
                self._graphs.clear()
                return self._eager_update(x, num_iters)
            # the captured kernels hold raw pointers into these caller-owned buffers: keep them alive with the graph
            keep = [q._prep_blob, q._ws] + [p.grad for p in q.parameters()]
            rec = (graph, static_x, losses, keep)
            self._graphs[key] = rec
        graph, static_x, losses, _ = rec
        static_x.copy_(x)
        graph.replay()
        self.quantizer._prep_key = None  # the parameters changed without their version counters moving
        # like the reference after its optim.zero_grad() (:681): no gradient is left on the parameters.  The graph
        # keeps writing into its own buffers (held by `keep`); an eager backward must never accumulate onto them.
        for p in self.quantizer.parameters():
            p.grad = None
        return losses

    def step(self, x: torch.Tensor) -> None:
        x = x.reshape(-1, self.quantizer.dim)
        num_iters = 2 if random.random() < self.two_iter_prob else 1
        diagnostics = self.cur_iter % 200 == 0
        graphed = self._use_graph and x.is_cuda and not diagnostics
        if diagnostics:
            # like the reference (:655-658): evaluated with the parameters this step starts from
            with torch.no_grad():
                det_losses = [float('%.3f' % self.quantizer.compute_loss(x, j)[0].item()) for j in range(6)]
        if graphed:
            (reconstruction_loss, logprob_loss, logits_entropy_loss,
             index_entropy_loss) = self._graphed_update(x.contiguous(), num_iters)
        else:
            (reconstruction_loss, logprob_loss, logits_entropy_loss,
             index_entropy_loss) = self._eager_update(x, num_iters)

        if diagnostics:
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

        self.scheduler.step()

        if self.cur_iter == self.phase_one_iters:
            self._begin_second_phase()
        self.cur_iter += 1

    def _init_optimizer(self):
        # same hyper-parameters as the reference (:722-727).  On a GPU: the fused multi-tensor implementation (one
        # kernel instead of ~12 foreach kernels of 10 us each; it is also the graph-capturable one)
        on_gpu = next(self.quantizer.parameters()).is_cuda
        self.optim = torch.optim.Adam(self.quantizer.parameters(), lr=self.lr, betas=(0.9, 0.98), eps=1e-9,
                                      weight_decay=1.0e-06,
                                      **({"fused": True, "capturable": self._use_graph} if on_gpu else {}))
        self._graphs.clear()
        self._warm.clear()
        step_size = (self.phase_one_iters if self.cur_iter == 0 else self.phase_two_iters) / 4
        self.scheduler = torch.optim.lr_scheduler.StepLR(self.optim, step_size=step_size, gamma=0.5)

    def _begin_second_phase(self):
        self.quantizer = self.quantizer.get_product_quantizer()
        self.lr *= 0.5
        self._init_optimizer()

    def get_quantizer(self) -> Quantizer:
        assert self.cur_iter >= self.phase_one_iters + self.phase_two_iters
        return self.quantizer
