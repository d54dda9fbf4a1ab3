"""Multi-GPU: replicas only.  The reference shards *prompts* across independent model replicas (Ray actors,
vispec/evaluation/gen_spec_answer_coco_caption.py:53-83); one generation (batch size 1) has no data to shard.
One process per GPU, no data-path collective; torch.distributed is used only to gather the results."""
from __future__ import annotations

from typing import List, Sequence


def shard_prompts(n_prompts: int, rank: int, world: int) -> range:
    """Contiguous chunks like the reference's `chunk_size = len(questions) // (num_gpus_total // num_gpus_per_model)`
    (gen_spec_answer_coco_caption.py:61-66), with the remainder spread over the first ranks so nothing is dropped."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, rem = divmod(n_prompts, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def gather_results(local: Sequence[dict], group=None) -> List[dict]:
    """All ranks' per-prompt records (question_id, new_tokens, wall_time, acceptance_length), ordered by question_id
    like `reorg_answer_file` (gen_spec_answer_coco_caption.py:288-299)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return sorted(local, key=lambda r: r["question_id"])
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, list(local), group=group)
    merged = [r for part in out for r in part]
    return sorted(merged, key=lambda r: r["question_id"])


def aggregate_throughput(records: Sequence[dict]) -> dict:
    """speed.py:46-98 reducers: mean tokens/s over prompts and mean accept length."""
    speeds = [sum(r["new_tokens"]) / sum(r["wall_time"]) for r in records if sum(r["wall_time"]) > 0]
    acc = [a for r in records for a in r.get("acceptance_length", [])]
    return {"mean_tokens_per_s": sum(speeds) / max(1, len(speeds)), "mean_accept_len": sum(acc) / max(1, len(acc)),
            "n_prompts": len(records)}
