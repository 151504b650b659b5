"""Layer distribution helper (a ``pass`` stub in the reference,
``parallelism/pipeline_parallel/utils.py:13-26``)."""
from typing import List


def distribute_layers(num_layers: int, num_stages: int) -> List[List[int]]:
    """Split ``num_layers`` layer indices over ``num_stages`` stages as evenly as
    possible; the remainder goes to the earliest stages."""
    if num_stages < 1:
        raise ValueError("num_stages must be >= 1")
    base, rem = divmod(num_layers, num_stages)
    out, start = [], 0
    for s in range(num_stages):
        n = base + (1 if s < rem else 0)
        out.append(list(range(start, start + n)))
        start += n
    return out
