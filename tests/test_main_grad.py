"""fp32 main-grad buckets for low-precision parameters (Megatron-style ``main_grad``): micro-batch gradients are
accumulated in fp32 bucket views — by post-accumulate hooks for ordinary gradients, inside the weight-gradient op for
linear layers (``fuse_wgrad_accumulation``) — and every consumer between backward and the optimizer step (global-norm
clipping, FusedAdamW) reads them through ``grad_of``.  CPU twin of the B200 path (``qn_gemm_bf16_wgrad_accum_f32``)."""
import copy

import pytest
import torch
import torch.nn as nn

from quintnet_b200 import ops
from quintnet_b200.optim import FusedAdamW, clip_grad_norm_global
from quintnet_b200.parallel.data_parallel.components.bucket import grad_of


class _Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.w1 = nn.Parameter(torch.randn(64, 32) * 0.2)
        self.b1 = nn.Parameter(torch.zeros(64))
        self.w2 = nn.Parameter(torch.randn(16, 64) * 0.2)
        self.ln = nn.LayerNorm(32)

    def forward(self, x):
        return ops.linear(ops.linear(self.ln(x), self.w1, self.b1, "gelu"), self.w2)


@pytest.mark.parametrize("fused", [False, True], ids=["hook-accumulate", "fused-wgrad-accumulate"])
def test_fp32_main_grads_accumulate_clip_and_step(fused):
    torch.manual_seed(0)
    ref = _Net()                                     # fp32 oracle
    net = copy.deepcopy(ref).to(torch.bfloat16)
    with torch.no_grad():                            # same (bf16-representable) weights on both sides
        for p, q in zip(ref.parameters(), net.parameters()):
            p.copy_(q.float())
    opt = FusedAdamW(net.parameters(), lr=1e-2, weight_decay=0.0, grad_dtype=torch.float32)
    prev = ops.fused_wgrad_accumulation()
    ops.set_fused_wgrad_accumulation(fused)
    try:
        M = 8
        xs = [torch.randn(24, 32) for _ in range(M)]
        opt.zero_grad()
        for x in xs:
            net(x.to(torch.bfloat16)).float().pow(2).mean().backward()
            ref(x.to(torch.bfloat16).float()).pow(2).mean().backward()
    finally:
        ops.set_fused_wgrad_accumulation(prev)
    for (n, p), q in zip(net.named_parameters(), ref.parameters()):
        assert p.grad is None, n                     # nothing is left (or accumulated) in bf16 .grad tensors
        assert p.main_grad.dtype == torch.float32 and grad_of(p) is p.main_grad
        torch.testing.assert_close(p.main_grad, q.grad, atol=2e-2, rtol=6e-2)
    # clipping scales the fp32 main grads (not the absent .grad tensors)
    before = [grad_of(p).clone() for p in net.parameters()]
    norm = clip_grad_norm_global(list(net.parameters()), 1e-3)
    assert norm is not None and norm > 1e-3
    for b, p in zip(before, net.parameters()):
        torch.testing.assert_close(grad_of(p), b * (1e-3 / (norm + 1e-6)), atol=1e-7, rtol=1e-4)
    w_before = net.w1.detach().clone()
    opt.step()
    assert not torch.equal(w_before, net.w1.detach())
    opt.zero_grad()
    assert all(float(p.main_grad.abs().sum()) == 0.0 and p.grad is None for p in net.parameters())


def test_main_grad_resolution_from_config():
    from quintnet_b200.coordinators.main_coordinator import resolve_main_grad_dtype
    assert resolve_main_grad_dtype({"dtype": "float32"}) is None
    assert resolve_main_grad_dtype({"dtype": "bfloat16"}) is torch.float32
    assert resolve_main_grad_dtype({"dtype": "bfloat16", "main_grad_dtype": "param"}) is None
