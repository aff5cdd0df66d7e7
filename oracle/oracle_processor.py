"""CPU attention processor with the REFERENCE's semantics, hosted on the Attention stand-in -- TEST INFRASTRUCTURE ONLY
(see paid_oracle.py): lets the UNet harness run a whole denoising loop on the CPU in fp32 exactly as the reference's
processors would (interpolation.py:573-679 outer, 707-804 inner, 581-584 deactivated), so the CUDA pipeline can be
compared end to end (oracle/gen_e2e_golden.py -> tests/golden/e2e_sd15_c1.npz)."""
from __future__ import annotations

import torch

import paid_oracle as O


class OracleAttnProcessor(torch.nn.Module):   # a Module, like the reference's processors (Attention registers it as a child)
    def __init__(self, mode: int = O.MODE_OUTER, is_fused: bool = True, size: int = 3, t: float = 0.5, rows: int = 512):
        super().__init__()
        self.mode, self.is_fused, self.size, self.rows = mode, is_fused, size, rows
        self.coef = torch.tensor([0.0, t, 1.0]) if size == 3 else torch.linspace(0, 1, size)
        self.activated = True

    # the state API the pipeline drives (reference interpolation.py:34-42 + the N-frame extension)
    def deactivate(self):
        self.activated = False

    def activate(self, t):
        assert 0 < t < 1
        self.activated, self.coef = True, torch.tensor([0.0, t, 1.0])

    def set_coefs(self, coef):
        coef = coef.detach().float().clone()
        coef[0], coef[-1] = 0, 1
        self.size, self.coef, self.activated = coef.numel(), coef, True

    def forward(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        assert attention_mask is None
        dt = hidden_states.dtype
        w = O.LayerWeights(attn.to_q.weight.to(dt), attn.to_k.weight.to(dt), attn.to_v.weight.to(dt),
                           attn.to_out[0].weight.to(dt), attn.to_out[0].bias.to(dt), attn.heads)
        if self.activated:
            return O.forward_chunked(hidden_states, encoder_hidden_states, w, self.coef.to(dt), self.mode, self.is_fused,
                                     attn.scale, rows=self.rows)
        zeros = torch.zeros(hidden_states.shape[0], dtype=dt)      # coefficients are unused in PLAIN mode
        return O.forward_chunked(hidden_states, encoder_hidden_states, w, zeros, O.MODE_PLAIN, False, attn.scale, rows=self.rows)
