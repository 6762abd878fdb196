"""Initial-noise draw with the semantics of diffusers' `randn_tensor` (DF/utils/torch_utils.py:152-199), which every reference
pipeline's `prepare_latents` uses: one generator or a LIST of per-sample generators (GRPOTrainer.evaluate passes per-prompt CPU generators,
grpo.py:110-113), CPU generators draw on the CPU and the tensor is moved afterwards, a CUDA generator cannot feed a CPU tensor."""
from __future__ import annotations

from typing import List, Optional, Sequence, Union

import torch


def randn_tensor(shape: Sequence[int], generator: Optional[Union[torch.Generator, List[torch.Generator]]] = None,
                 device: Optional[Union[str, torch.device]] = None, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    device = torch.device(device) if device is not None else torch.device("cpu")
    shape = tuple(shape)
    rand_device = device
    if generator is not None:
        gen_type = (generator[0] if isinstance(generator, list) else generator).device.type
        if gen_type != device.type and gen_type == "cpu":
            rand_device = torch.device("cpu")
        elif gen_type != device.type and gen_type == "cuda":
            raise ValueError(f"Cannot generate a {device} tensor from a generator of type {gen_type}.")
    if isinstance(generator, list) and len(generator) == 1:
        generator = generator[0]
    if isinstance(generator, list):
        one = (1,) + shape[1:]
        parts = [torch.randn(one, generator=generator[i], device=rand_device, dtype=dtype) for i in range(shape[0])]
        return torch.cat(parts, dim=0).to(device)
    return torch.randn(shape, generator=generator, device=rand_device, dtype=dtype).to(device)
