"""Result records and the speed / accept-length reducer of the reference's evaluation scripts,
so that numbers produced with this package are comparable column for column with the reference's README table.

* answer records: `vispec/evaluation/gen_spec_answer_coco_caption.py:206-285` (per-question loop, JSONL schema
  `question_id, answer_id, model_id, choices[{index, turns, idxs, new_tokens, wall_time, acceptance_length}], tstamp`)
  and `:288-299` (`reorg_answer_file`);
* reducer: `vispec/evaluation/speed.py:46-98` (mean over prompts of tokens / wall time for the speculative run, the
  same for the baseline run with the tokens re-counted from the decoded text, their ratio, and the mean accept length).

Host-side Python only: nothing here touches the decode path.  The dataset prompt builders and the Ray launcher stay
out of scope; `replicas.py` shards prompts over one process per GPU instead."""
from __future__ import annotations

import json
import os
import time
import uuid
from typing import Callable, Dict, Iterable, List, Optional, Sequence

import torch

CHOICE_KEYS = ("index", "turns", "idxs", "new_tokens", "wall_time", "acceptance_length")
ANSWER_KEYS = ("question_id", "answer_id", "model_id", "choices", "tstamp")


def _sync():
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def clean_output(text: str, special_tokens_map: Optional[Dict] = None) -> str:
    """gen_spec_answer_coco_caption.py:238-249: drop the special tokens, strip, drop a leading "Assistant:"."""
    for special_token in (special_tokens_map or {}).values():
        for tok in (special_token if isinstance(special_token, list) else [special_token]):
            text = text.replace(tok, "")
    text = text.strip()
    if text.startswith("Assistant:"):
        text = text.replace("Assistant:", "", 1).strip()
    return text


def run_question(model, model_inputs: Dict, temperature: float = 0.0, num_choices: int = 1,
                 decode: Optional[Callable[[torch.Tensor], str]] = None, **generate_kwargs) -> List[dict]:
    """One question through `specgenerate`, `num_choices` times, exactly as the reference's loop does
    (gen_spec_answer_coco_caption.py:206-273): seed = choice index, wall time around the whole call between two device
    synchronisations, the prompt stripped from the output, ids beyond the vocabulary mapped to 0.  `decode` turns the
    new ids into text (default: the model's tokenizer if it can decode, else the ids joined by spaces)."""
    tokenizer = model.get_tokenizer()
    if decode is None:
        if hasattr(tokenizer, "decode"):
            def decode(ids):
                return clean_output(tokenizer.decode(ids, spaces_between_special_tokens=False),
                                    getattr(tokenizer, "special_tokens_map", None))
        else:
            def decode(ids):
                return " ".join(str(int(t)) for t in ids)
    vocab = getattr(tokenizer, "vocab_size", None) or model.vocab_size
    choices = []
    for i in range(num_choices):
        torch.manual_seed(i)
        _sync()
        start = time.time()
        output_ids, new_token, idx, accp_len = model.specgenerate(**model_inputs, temperature=temperature, log=True,
                                                                  return_acceptance_len=True, **generate_kwargs)
        _sync()
        total = time.time() - start
        new_ids = output_ids[0][len(model_inputs["input_ids"][0]):].clone()
        new_ids[new_ids > vocab] = 0
        choices.append({"index": i, "turns": [decode(new_ids)], "idxs": [int(idx)], "new_tokens": [int(new_token)],
                        "wall_time": [total], "acceptance_length": [int(a) for a in accp_len]})
    return choices


def answer_record(question_id, model_id: str, choices: Sequence[dict]) -> dict:
    """gen_spec_answer_coco_caption.py:277-284 (shortuuid is replaced by 22 hex digits of a uuid4)."""
    for c in choices:
        missing = [k for k in CHOICE_KEYS if k not in c]
        if missing:
            raise ValueError(f"choice record lacks {missing}")
    return {"question_id": question_id, "answer_id": uuid.uuid4().hex[:22], "model_id": model_id,
            "choices": list(choices), "tstamp": time.time()}


def append_answer(answer_file: str, record: dict) -> None:
    """gen_spec_answer_coco_caption.py:275-285: one JSON object per line, appended."""
    answer_file = os.path.expanduser(answer_file)
    d = os.path.dirname(answer_file)
    if d:
        os.makedirs(d, exist_ok=True)
    with open(answer_file, "a") as fout:
        fout.write(json.dumps(record) + "\n")


def reorg_answer_file(answer_file: str) -> None:
    """gen_spec_answer_coco_caption.py:288-299: sort by question id, keep the last record of a repeated id."""
    answers = {}
    with open(answer_file, "r") as fin:
        for line in fin:
            if line.strip():
                answers[json.loads(line)["question_id"]] = line if line.endswith("\n") else line + "\n"
    with open(answer_file, "w") as fout:
        for qid in sorted(answers.keys()):
            fout.write(answers[qid])


def read_answers(answer_file: str) -> List[dict]:
    with open(answer_file, "r", encoding="utf-8") as f:
        return [json.loads(line) for line in f if line.strip()]


def speed_report(spec_records: Iterable[dict], base_records: Optional[Iterable[dict]] = None,
                 count_tokens: Optional[Callable[[str], int]] = None) -> dict:
    """speed.py:46-98.  Speculative run: per prompt sum(new_tokens) / sum(wall_time) of choice 0; accept lengths pooled
    over prompts.  Baseline run: tokens re-counted from the decoded turns with `count_tokens` (the reference uses
    `len(tokenizer(turn).input_ids) - 1`), falling back to the recorded `new_tokens` when no counter is given."""
    speeds, acc_len, new_tokens = [], [], []
    for dp in spec_records:
        c = dp["choices"][0]
        tokens, times = sum(c["new_tokens"]), sum(c["wall_time"])
        acc_len += c["acceptance_length"]
        speeds.append(tokens / times)
        new_tokens.append(tokens)
    out = {"n_prompts": len(speeds),
           "acc_len": sum(acc_len) / len(acc_len) if acc_len else 0.0,
           "new_tok": sum(new_tokens) / len(new_tokens) if new_tokens else 0.0,
           "speed": sum(speeds) / len(speeds) if speeds else 0.0}
    if base_records is not None:
        speeds0 = []
        for dp in base_records:
            c = dp["choices"][0]
            if count_tokens is not None:
                tokens = sum(count_tokens(t) for t in c["turns"])
            else:
                tokens = sum(c["new_tokens"])
            speeds0.append(tokens / sum(c["wall_time"]))
        out["speed0"] = sum(speeds0) / len(speeds0) if speeds0 else 0.0
        out["ratio"] = out["speed"] / out["speed0"] if out["speed0"] else float("nan")
    return out


def main(argv=None) -> int:
    import argparse

    ap = argparse.ArgumentParser(description="accept length / tokens per second / speed-up from answer files")
    ap.add_argument("--spec", required=True, help="JSONL written by the speculative run")
    ap.add_argument("--base", default=None, help="JSONL written by the baseline (autoregressive) run")
    ap.add_argument("--tokenizer", default=None, help="HF tokenizer path used to re-count the baseline's tokens")
    args = ap.parse_args(argv)
    counter = None
    if args.tokenizer:
        from transformers import AutoTokenizer

        tok = AutoTokenizer.from_pretrained(args.tokenizer)

        def counter(text):
            return len(tok(text).input_ids) - 1
    rep = speed_report(read_answers(args.spec), read_answers(args.base) if args.base else None, counter)
    print(json.dumps(rep))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
