"""Row a-L1 (CPU): the ChatML templates and `tokenizer_image_token` against golden output of the reference's own
longva/conversation.py (:414-423, 525-562, 587-590) and longva/mm_utils.py (:341-360) —
tests/golden/prompts_ref.json, written by tests/golden/make_golden_prompts.py."""
import json
import os

import pytest
import torch

from fakes import FakeTokenizer
from streamchat_b200.longva import constants as K
from streamchat_b200.longva.conversation import SeparatorStyle, conv_templates
from streamchat_b200.longva.mm_utils import tokenizer_image_token


class BosTokenizer(FakeTokenizer):
    bos_token_id = 1

    def __call__(self, text, **kw):
        enc = super().__call__(text, **kw)
        enc["input_ids"] = [1] + list(enc["input_ids"])
        return enc


@pytest.fixture(scope="module")
def golden(golden_dir):
    return json.load(open(os.path.join(golden_dir, "prompts_ref.json")))


MESSAGES = {
    "open_turn": [("user", "<image>\nWhat is on the table?"), ("assistant", None)],
    "two_rounds": [("user", "<image>\nDescribe the room."), ("assistant", "A kitchen with a red kettle."),
                   ("user", "Where did I put the keys?"), ("assistant", None)],
    "no_image": [("user", "You need to write a summary of the following."), ("assistant", None)],
}


def test_constants(golden):
    for k, v in golden["constants"].items():
        assert getattr(K, k) == v, k


@pytest.mark.parametrize("name", ["qwen_1_5", "qwen_1_5_ego", "qwen_1_5_caption", "qwen_1_5_summarize"])
def test_templates_equal_reference(golden, name):
    ref = golden["templates"][name]
    t = conv_templates[name]
    assert t.system == ref["system"] and list(t.roles) == ref["roles"] and t.sep == ref["sep"]
    assert t.sep_style == SeparatorStyle[ref["sep_style"]] and t.version == ref["version"]
    for case, msgs in MESSAGES.items():
        c = t.copy()
        for role, text in msgs:
            c.append_message(c.roles[0] if role == "user" else c.roles[1], text)
        assert c.get_prompt() == ref["prompts"][case], (name, case)
    assert t.messages == []                      # copy() does not alias the template's message list


@pytest.mark.parametrize("label,tok", [("no_bos", FakeTokenizer()), ("bos", BosTokenizer())])
def test_tokenizer_image_token_equals_reference(golden, label, tok):
    for prompt, ref in golden["tokenizer_image_token"][label].items():
        assert tokenizer_image_token(prompt, tok, K.IMAGE_TOKEN_INDEX) == ref["list"], (label, prompt)
        pt = tokenizer_image_token(prompt, tok, K.IMAGE_TOKEN_INDEX, return_tensors="pt")
        assert pt.dtype == torch.long and pt.tolist() == ref["pt"], (label, prompt)
    with pytest.raises(ValueError):
        tokenizer_image_token("x", tok, K.IMAGE_TOKEN_INDEX, return_tensors="np")
