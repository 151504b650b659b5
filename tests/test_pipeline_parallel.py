"""Pipeline-parallel building blocks on CPU / gloo (reference ``tests/test_pipeline_parallel.py``:
``test_layer_distribution``, ``test_pipeline_forward_pass`` — there they need GPUs)."""
import pytest
import torch
import torch.distributed as dist

from tests._dist import run_distributed
from quintnet_b200 import PipelineParallelWrapper, init_process_groups
from quintnet_b200.core.communication import drain_pipeline_sends, pipeline_communicate
from quintnet_b200.models import Model
from quintnet_b200.parallel.pipeline_parallel.utils import balanced_gpt2_split, distribute_layers


def test_layer_distribution():
    assert distribute_layers(8, 2) == [[0, 1, 2, 3], [4, 5, 6, 7]]
    assert distribute_layers(7, 3) == [[0, 1, 2], [3, 4], [5, 6]]              # remainder to the earliest stages
    assert distribute_layers(4, 1) == [[0, 1, 2, 3]]
    assert distribute_layers(24, 2, counts=[13, 11])[1][0] == 13
    with pytest.raises(ValueError):
        distribute_layers(8, 2, counts=[5, 4])
    for n, pp in ((24, 2), (24, 4), (12, 2), (36, 4)):
        counts = balanced_gpt2_split(n, pp)
        assert sum(counts) == n and len(counts) == pp and counts[-1] <= counts[0] and min(counts) >= 1


def _forward_backward(rank, world):
    """A ViT split over the pipeline ranks computes the same logits and input-side gradients as the whole model."""
    pg = init_process_groups("cpu", [world], ["pp"])
    torch.manual_seed(0)
    full = Model(hidden_dim=32, depth=4, n_heads=4)
    ref = Model(hidden_dim=32, depth=4, n_heads=4)
    ref.load_state_dict(full.state_dict())
    stage = PipelineParallelWrapper(full, pg.device_mesh, rank, pg.get_group("pp"), world, torch.device("cpu"))
    assert stage.is_first_stage == (rank == 0) and stage.is_last_stage == (rank == world - 1)
    n_blocks = sum(1 for m in stage.local_module if m.__class__.__name__ == "TransformerBlock")
    assert n_blocks == len(distribute_layers(4, world)[rank])
    x = torch.randn(4, 1, 28, 28, generator=torch.Generator().manual_seed(1))
    shapes = stage.get_tensor_shapes(4)
    kw = dict(pp_group=pg.get_group("pp"), pp_rank=rank, device=torch.device("cpu"), dtype=torch.float32,
              is_first_stage=stage.is_first_stage, is_last_stage=stage.is_last_stage, shapes=shapes)
    inp = pipeline_communicate("recv_forward", **kw)
    out = stage.forward(x if stage.is_first_stage else inp)
    pipeline_communicate("send_forward", tensor=out, **kw)
    expect = ref(x)
    if stage.is_last_stage:
        assert torch.allclose(out, expect, atol=1e-5)
        loss = out.square().mean()
        gin = stage.backward(inp, loss, None)
    else:
        g = pipeline_communicate("recv_backward", **kw)
        gin = stage.backward(inp, out, g)
    pipeline_communicate("send_backward", tensor=gin, **kw)
    drain_pipeline_sends()
    expect.square().mean().backward()
    if stage.is_first_stage:
        assert torch.allclose(full.embedding.pos_embed.grad, ref.embedding.pos_embed.grad, atol=1e-5)
    if stage.is_last_stage:
        assert torch.allclose(full.classification_head.fc.weight.grad if hasattr(full.classification_head, "fc")
                              else next(full.classification_head.parameters()).grad,
                              ref.classification_head.fc.weight.grad if hasattr(ref.classification_head, "fc")
                              else next(ref.classification_head.parameters()).grad, atol=1e-5)
    dist.barrier()


@pytest.mark.parametrize("world", [2, 4])
def test_pipeline_forward_pass(world):
    run_distributed(_forward_backward, world)
