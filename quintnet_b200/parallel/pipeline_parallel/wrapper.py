"""``PipelineParallelWrapper`` — turns a model into this rank's pipeline stage.

Contract parity with the reference (``parallelism/pipeline_parallel/wrapper.py:33-249``):
ctor ``(model, device_mesh, rank, pp_group, pp_size, device, stage_module=None)``,
attributes ``rank/world_size/group/is_first_stage/is_last_stage/local_module``,
``get_tensor_shapes(batch)``, ``forward(x, **kw)`` and the custom
``backward(input_tensor, output_tensor, output_tensor_grad)`` the schedules call.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from .utils import distribute_layers


class PipelineParallelWrapper(nn.Module):
    def __init__(self, model: Optional[nn.Module], device_mesh, rank: int, pp_group, pp_size: int,
                 device, stage_module: Optional[nn.Module] = None):
        super().__init__()
        self.device_mesh = device_mesh
        self.rank, self.world_size, self.group = rank, pp_size, pp_group
        self.is_first_stage = rank == 0
        self.is_last_stage = rank == pp_size - 1
        self.device = device
        self.tensor_shapes = None
        if stage_module is not None:
            self.local_module = stage_module.to(device)
            self.depth = None
            self.layer_distribution = None
        else:
            if not hasattr(model, "blocks"):
                raise ValueError("PipelineParallelWrapper needs a model with .embedding/.blocks/"
                                 ".classification_head (or pass stage_module=)")
            self.depth = len(model.blocks)
            self.layer_distribution = self._distribute_layers(self.depth)
            self.tensor_shapes = self._infer_tensor_shapes(model)
            self.local_module = self._build_local_module(model).to(device)

    # ------------------------------------------------------------ stage build
    def _distribute_layers(self, num_layers: int) -> List[int]:
        return distribute_layers(num_layers, self.world_size)[self.rank]

    @staticmethod
    def _infer_tensor_shapes(model):
        emb = getattr(model, "embedding", None)
        if emb is not None and hasattr(emb, "pos_embed"):
            return (int(emb.pos_embed.shape[1]), int(emb.pos_embed.shape[2]))
        if hasattr(model, "get_tensor_shapes"):
            s = model.get_tensor_shapes(1)
            return tuple(int(v) for v in s[1:])
        return None

    def _build_local_module(self, model) -> nn.Module:
        layers: List[nn.Module] = []
        if self.is_first_stage and getattr(model, "embedding", None) is not None:
            layers.append(model.embedding)
        layers.extend(model.blocks[i] for i in self.layer_distribution)
        if self.is_last_stage:
            head = getattr(model, "classification_head", None)
            if head is None:
                head = getattr(model, "head", None)
            if head is not None:
                layers.append(head)
        return nn.Sequential(*layers)

    def get_tensor_shapes(self, batch_size: int):
        if hasattr(self.local_module, "get_tensor_shapes"):
            return tuple(self.local_module.get_tensor_shapes(batch_size))
        if self.tensor_shapes is None:
            raise RuntimeError("tensor shapes unknown for this stage")
        return (batch_size,) + tuple(self.tensor_shapes)

    # ------------------------------------------------------------ CUDA graphs
    def enable_cuda_graphs(self, input_shape, dtype, input_is_float: bool = True) -> bool:
        """Capture this stage's micro-batch forward AND backward into CUDA graphs (``torch.cuda.make_graphed_callables``):
        the latency-bound regime (ViT-MNIST: a stage is ~150 kernels of a few microseconds each) is bound by launch
        overhead, and a replay is two launches.  Every kernel of this library is capture-safe (raw launches on the
        current stream, tensor maps / descriptors baked at capture, device-side launch counters in the one-shot
        all-reduce), and NCCL collectives record into the graph.  Shapes become static: only micro-batches of exactly
        ``input_shape`` take the graphed path (others run eagerly).  Returns False (and stays eager) if capture fails."""
        if not (torch.cuda.is_available() and str(self.device).startswith("cuda")):
            return False
        try:
            sample = torch.randn(tuple(input_shape), device=self.device, dtype=dtype)
            if not self.is_first_stage and input_is_float:
                sample.requires_grad_(True)
            eager = self.local_module
            was_training = eager.training
            eager.train()
            graphed = torch.cuda.make_graphed_callables(eager, (sample,), num_warmup_iters=3, allow_unused_input=True)
            eager.train(was_training)
            self.local_module = graphed
            self._graphed_shape = tuple(input_shape)
            for p in self.local_module.parameters():      # the warm-up / capture backward passes left garbage gradients
                p.grad = None
            return True
        except Exception as e:  # noqa: BLE001 — capture is an optimisation; report and stay eager
            print(f"[quintnet_b200] CUDA-graph capture of pipeline stage {self.rank} failed, running eagerly: "
                  f"{type(e).__name__}: {str(e)[:300]}", flush=True)
            return False

    # -------------------------------------------------------------- execution
    def forward(self, x, **kwargs):
        if kwargs:
            return self.local_module(x, **kwargs)
        return self.local_module(x)

    def backward(self, input_tensor, output_tensor, output_tensor_grad):
        """Backward of one micro-batch through this stage.

        Returns the gradient w.r.t. the stage input (``None`` on the first
        stage).  On the last stage ``output_tensor`` is the (already scaled)
        scalar loss and ``output_tensor_grad`` is ``None``.
        """
        if input_tensor is not None and input_tensor.requires_grad:
            input_tensor.retain_grad()
        if output_tensor_grad is None and output_tensor.numel() == 1:
            torch.autograd.backward(output_tensor)
        else:
            if output_tensor_grad is None:
                output_tensor_grad = torch.ones_like(output_tensor)
            torch.autograd.backward(output_tensor, grad_tensors=output_tensor_grad)
        grad = None
        if input_tensor is not None:
            grad = input_tensor.grad
            input_tensor.grad = None
        return grad
