"""Nucleus sampling of the next token, as the reference does it (moondream/torch/moondream.py:270-278 `_apply_top_p`,
:312-318 and :524-530 `softmax(logits / temperature) -> top-p -> torch.multinomial`).

Interim host implementation: the engine hands over the step's bf16 logits (device -> host, [batch, vocab]) and the
arithmetic below runs with the same torch ops, dtype and global RNG stream as the reference on CPU, so identical
logits and an identical `torch.manual_seed` give identical tokens.  It costs one synchronisation and a vocabulary
sort per token; the on-device kernel replaces it without changing callers.
"""
from __future__ import annotations

from typing import Optional

import torch


def apply_top_p(probs: torch.Tensor, top_p: float) -> torch.Tensor:
    """Zero everything outside the smallest prefix of the sorted distribution whose mass exceeds `top_p`
    (the token that crosses the threshold is kept), renormalise, scatter back (moondream.py:270-278)."""
    ordered, order = torch.sort(probs, dim=-1, descending=True)
    mass_before = torch.cumsum(ordered, dim=-1) - ordered          # same expression, same dtype as the reference
    ordered = ordered.masked_fill(mass_before > top_p, 0.0)
    ordered = ordered / ordered.sum(dim=-1, keepdim=True)
    return torch.zeros_like(probs).scatter_(-1, order, ordered)


def sample_next(logits: torch.Tensor, temperature: float, top_p: float,
                generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """logits [batch, vocab] (any float dtype, CPU) -> int64 [batch] token ids.  temperature == 0 is argmax
    (moondream.py:312-313); otherwise the arithmetic stays in the logits' dtype like the reference's."""
    if temperature == 0:
        return torch.argmax(logits, dim=-1)
    probs = torch.softmax(logits / temperature, dim=-1)
    probs = apply_top_p(probs, top_p)
    return torch.multinomial(probs, num_samples=1, generator=generator).squeeze(1)


class HostSampler:
    """Callable the engine's decode loop invokes once per step with the device logits of that step."""

    def __init__(self, temperature: float, top_p: float, generator: Optional[torch.Generator] = None):
        if temperature < 0:
            raise ValueError("temperature must be >= 0")
        self.temperature = float(temperature)
        self.top_p = float(top_p)
        self.generator = generator
        self._host: Optional[torch.Tensor] = None

    def __call__(self, logits: torch.Tensor) -> torch.Tensor:
        if logits.device.type != "cpu":
            if self._host is None or self._host.shape != logits.shape or self._host.dtype != logits.dtype:
                self._host = torch.empty(logits.shape, dtype=logits.dtype, pin_memory=True)
            self._host.copy_(logits, non_blocking=False)          # the per-token synchronisation
            logits = self._host
        return sample_next(logits, self.temperature, self.top_p, self.generator).to(torch.int32)
