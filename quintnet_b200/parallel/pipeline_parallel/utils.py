"""Layer distribution helper (a ``pass`` stub in the reference,
``parallelism/pipeline_parallel/utils.py:13-26``)."""
from typing import List


from typing import Optional, Sequence


def distribute_layers(total_layers: int, num_stages: int,
                      counts: Optional[Sequence[int]] = None) -> List[List[int]]:
    """Split ``total_layers`` layer indices over ``num_stages`` stages.  Default: as evenly as possible,
    remainder to the earliest stages (the reference's rule).  ``counts`` gives an explicit number of
    layers per stage (e.g. ``[13, 11]`` to offset the lm_head + loss work of the last GPT-2 stage)."""
    if num_stages < 1:
        raise ValueError("num_stages must be >= 1")
    if counts is not None:
        counts = [int(c) for c in counts]
        if len(counts) != num_stages or sum(counts) != total_layers or min(counts) < 0:
            raise ValueError(f"pp layer split {counts} does not cover {total_layers} layers over {num_stages} stages")
    else:
        base, rem = divmod(total_layers, num_stages)
        counts = [base + (1 if s < rem else 0) for s in range(num_stages)]
    out, start = [], 0
    for n in counts:
        out.append(list(range(start, start + n)))
        start += n
    return out


def balanced_gpt2_split(n_layer: int, pp_size: int, head_cost_in_blocks: float = 2.5) -> List[int]:
    """Layer counts that equalise stage time when the last stage also runs the lm_head + cross-entropy
    (worth ``head_cost_in_blocks`` transformer blocks of forward+backward time)."""
    if pp_size == 1:
        return [n_layer]
    per = (n_layer + head_cost_in_blocks) / pp_size
    last = max(1, int(round(per - head_cost_in_blocks)))
    rest = n_layer - last
    if rest < pp_size - 1:
        return [len(x) for x in distribute_layers(n_layer, pp_size)]
    base, rem = divmod(rest, pp_size - 1)
    return [base + (1 if s < rem else 0) for s in range(pp_size - 1)] + [last]
