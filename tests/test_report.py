"""Result records + speed reducer (vispec_b200/report.py) against the reference's schema and arithmetic
(vispec/evaluation/gen_spec_answer_coco_caption.py:206-299, speed.py:46-98)."""
import json
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vispec_b200 import report  # noqa: E402


def _choice(new_tokens, wall, acc, turn="a b c"):
    return {"index": 0, "turns": [turn], "idxs": [len(acc)], "new_tokens": [new_tokens], "wall_time": [wall],
            "acceptance_length": acc}


def test_answer_schema_matches_reference(tmp_path):
    rec = report.answer_record(17, "llava-vispec", [_choice(12, 0.5, [2, 3, 1])])
    assert tuple(rec.keys()) == report.ANSWER_KEYS == ("question_id", "answer_id", "model_id", "choices", "tstamp")
    assert tuple(rec["choices"][0].keys()) == report.CHOICE_KEYS
    assert len(rec["answer_id"]) == 22
    with pytest.raises(ValueError):
        report.answer_record(1, "m", [{"index": 0}])
    f = tmp_path / "sub" / "answers.jsonl"
    report.append_answer(str(f), rec)
    report.append_answer(str(f), rec)
    lines = f.read_text().splitlines()
    assert len(lines) == 2 and json.loads(lines[0])["question_id"] == 17


def test_reorg_sorts_and_keeps_last_duplicate(tmp_path):
    f = tmp_path / "a.jsonl"
    for qid, mid in ((5, "x"), (2, "x"), (5, "y"), (1, "x")):
        report.append_answer(str(f), report.answer_record(qid, mid, [_choice(4, 1.0, [1])]))
    report.reorg_answer_file(str(f))
    recs = report.read_answers(str(f))
    assert [r["question_id"] for r in recs] == [1, 2, 5]
    assert recs[2]["model_id"] == "y"


def test_speed_report_arithmetic():
    spec = [report.answer_record(0, "m", [_choice(30, 0.10, [2, 2, 3])]),
            report.answer_record(1, "m", [_choice(20, 0.20, [0, 1])])]
    base = [report.answer_record(0, "b", [_choice(30, 0.30, [], turn="w " * 30)]),
            report.answer_record(1, "b", [_choice(20, 0.40, [], turn="w " * 20)])]
    rep = report.speed_report(spec, base, count_tokens=lambda t: len(t.split()))
    assert rep["n_prompts"] == 2
    assert rep["acc_len"] == pytest.approx(8 / 5)
    assert rep["new_tok"] == pytest.approx(25.0)
    assert rep["speed"] == pytest.approx((300.0 + 100.0) / 2)
    assert rep["speed0"] == pytest.approx((100.0 + 50.0) / 2)
    assert rep["ratio"] == pytest.approx(200.0 / 75.0)
    # without a token counter the recorded new_tokens of the baseline are used
    assert report.speed_report(spec, base)["speed0"] == pytest.approx(75.0)
    assert "ratio" not in report.speed_report(spec)


def test_clean_output():
    sm = {"bos_token": "<s>", "additional_special_tokens": ["<image>", "<pad>"]}
    assert report.clean_output("<s> Assistant: a cat<pad><pad> ", sm) == "a cat"


def test_cli(tmp_path, capsys):
    f = tmp_path / "s.jsonl"
    report.append_answer(str(f), report.answer_record(0, "m", [_choice(10, 0.5, [1, 1])]))
    assert report.main(["--spec", str(f), "--base", str(f)]) == 0
    out = json.loads(capsys.readouterr().out)
    assert out["speed"] == pytest.approx(20.0) and out["ratio"] == pytest.approx(1.0)


@pytest.mark.gpu
def test_run_question_records_a_generation(tmp_path):
    from test_gpu_e2e import build, load_case
    g, tcfg, dcfg, tw, dw, ids, emb, mask = load_case("bigram_noisy_img")
    model = build(tcfg, dcfg, tw, dw)
    inputs = {"input_ids": ids.to("cuda"), "inputs_embeds": emb.to("cuda"), "image_mask": mask.to("cuda")}
    choices = report.run_question(model, inputs, temperature=0.0, num_choices=2, max_new_tokens=24)
    assert len(choices) == 2 and [c["index"] for c in choices] == [0, 1]
    c = choices[0]
    assert c["new_tokens"][0] == sum(c["acceptance_length"]) + len(c["acceptance_length"])
    assert c["turns"] == choices[1]["turns"]                       # greedy: both choices identical
    assert c["wall_time"][0] > 0 and c["idxs"][0] == len(c["acceptance_length"]) - 1
    rec = report.answer_record(3, "tiny", choices)
    f = tmp_path / "a.jsonl"
    report.append_answer(str(f), rec)
    rep = report.speed_report(report.read_answers(str(f)))
    assert rep["new_tok"] == c["new_tokens"][0] and rep["acc_len"] > 0
