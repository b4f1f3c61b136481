"""A deterministic stand-in for the reference's AlphaTensor network (model.py is out of scope and is not
available on the GPU box): fwd_infer returns seeded random sparse actions and a scalar value, so the MCTS
of act.py can be compared call for call between the reference and this package."""
import torch


class FakeAlphaTensor:
    def __init__(self, dim_3d=4, n_samples=4, n_logits=3, seed=0):
        self.dim_3d, self.n_samples, self.n_logits = dim_3d, n_samples, n_logits
        self.n_steps = 3 * dim_3d
        self.device = torch.device("cpu")
        self.gen = torch.Generator().manual_seed(seed)
        self.calls = 0

    def fwd_infer(self, state, scalars):
        self.calls += 1
        aa = torch.randint(0, self.n_logits, (1, self.n_samples, self.n_steps), generator=self.gen)
        sparse = torch.rand(1, self.n_samples, self.n_steps, generator=self.gen) < 0.5
        aa[sparse] = 1  # token 1 == coefficient 0 under act.py's fixed shift of 1
        pp = torch.full((1, self.n_samples), 1.0 / self.n_samples)
        qq = -3.0 * torch.rand(1, generator=self.gen)
        return aa, pp, qq
