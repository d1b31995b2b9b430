"""torch.optim.Adam-compatible optimizer whose step() is one launch of our fused kernel.

Replaces optim.Adam(..., lr, betas=[0.5, 0.999]) at train/dcgan_trainer.py:61-62 and its .step()
(:180,189).  It subclasses torch.optim.Adam so param_groups / state_dict() / load_state_dict() keep the
reference's checkpoint format ('state': {idx: {step, exp_avg, exp_avg_sq}}, 'param_groups')."""
import torch

from .. import ops


class FusedAdam(torch.optim.Adam):
    def __init__(self, params, lr, betas, flat, eps=1e-8):
        super().__init__(params, lr=lr, betas=tuple(betas), eps=eps)
        self.flat = flat
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=flat.flat.device)
        self.steps_done = 0

    @torch.no_grad()
    def step(self, closure=None):
        g = self.param_groups[0]
        b1, b2 = g['betas']
        ops.adam(self.flat.flat, self.flat.grad, self.flat.exp_avg, self.flat.exp_avg_sq, float(g['lr']),
                 float(b1), float(b2), float(g['eps']), self.step_dev)
        ops.adam_advance(self.step_dev)
        self.steps_done += 1

    def _publish_state(self):
        if self.steps_done == 0:
            return
        m, v = self.flat.views(self.flat.exp_avg), self.flat.views(self.flat.exp_avg_sq)
        for p, mi, vi in zip(self.flat.params, m, v):
            self.state[p] = {'step': torch.tensor(float(self.steps_done)), 'exp_avg': mi, 'exp_avg_sq': vi}

    def state_dict(self):
        self._publish_state()
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        steps = 0
        for (o, k), p in zip(self.flat.offsets, self.flat.params):
            st = self.state.get(p)
            if st:
                self.flat.exp_avg[o:o + k].copy_(st['exp_avg'].reshape(-1))
                self.flat.exp_avg_sq[o:o + k].copy_(st['exp_avg_sq'].reshape(-1))
                steps = int(float(st['step']))
        self.steps_done = steps
        self.step_dev.fill_(steps)
