"""Golden prompts / token lists for row a-L1, produced by the REFERENCE's own `longva.conversation` templates
(conv_templates[...] :414-423, 525-562, 587-590) and `longva.mm_utils.tokenizer_image_token` (:341-360).
Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_prompts.py

Writes prompts_ref.json.  The tokenizers are the deterministic stand-ins of tests/fakes.py (with and without a BOS).
"""
import importlib.util
import json
import os
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.dirname(HERE))
from fakes import FakeTokenizer  # noqa: E402


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


class BosTokenizer(FakeTokenizer):
    bos_token_id = 1

    def __call__(self, text, **kw):
        enc = super().__call__(text, **kw)
        enc["input_ids"] = [1] + list(enc["input_ids"])
        return enc


MESSAGES = {
    "open_turn": [("user", "<image>\nWhat is on the table?"), ("assistant", None)],
    "two_rounds": [("user", "<image>\nDescribe the room."), ("assistant", "A kitchen with a red kettle."),
                   ("user", "Where did I put the keys?"), ("assistant", None)],
    "no_image": [("user", "You need to write a summary of the following."), ("assistant", None)],
}
TOKEN_CASES = ["<image>\nWhat is this?", "before <image> after", "<image>", "no image here", "a <image> b <image> c",
               "<image><image> twice", "trailing <image>", ""]


def main():
    # conversation.py and mm_utils.py loaded as plain files (importing the `longva` package drags the model zoo in)
    sys.modules.setdefault("longva", type(sys)("longva"))
    const = _load("longva.constants", os.path.join(REF, "longva", "constants.py"))
    sys.modules["longva.constants"] = const
    conv = _load("longva.conversation", os.path.join(REF, "longva", "conversation.py"))
    src = open(os.path.join(REF, "longva", "mm_utils.py")).read()
    import ast
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "tokenizer_image_token")
    ns = {"IMAGE_TOKEN_INDEX": const.IMAGE_TOKEN_INDEX}
    import torch
    ns["torch"] = torch
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "mm_utils.py", "exec"), ns)
    tok_fn = ns["tokenizer_image_token"]

    out = {"templates": {}, "tokenizer_image_token": {}, "constants": {
        k: getattr(const, k) for k in ("IMAGE_TOKEN_INDEX", "DEFAULT_IMAGE_TOKEN", "DEFAULT_IM_START_TOKEN",
                                       "DEFAULT_IM_END_TOKEN", "DEFAULT_IMAGE_PATCH_TOKEN", "IGNORE_INDEX")}}
    for name in ("qwen_1_5", "qwen_1_5_ego", "qwen_1_5_caption", "qwen_1_5_summarize"):
        t = conv.conv_templates[name]
        entry = {"system": t.system, "roles": list(t.roles), "sep": t.sep, "sep_style": t.sep_style.name,
                 "version": t.version, "prompts": {}}
        for case, msgs in MESSAGES.items():
            c = t.copy()
            for role, text in msgs:
                c.append_message(c.roles[0] if role == "user" else c.roles[1], text)
            entry["prompts"][case] = c.get_prompt()
            assert t.messages == [] or len(t.messages) == len(conv.conv_templates[name].messages)   # copy() did not alias
        out["templates"][name] = entry
    for label, tk in (("no_bos", FakeTokenizer()), ("bos", BosTokenizer())):
        cases = {}
        for p in TOKEN_CASES:
            cases[p] = {"list": tok_fn(p, tk, const.IMAGE_TOKEN_INDEX),
                        "pt": tok_fn(p, tk, const.IMAGE_TOKEN_INDEX, return_tensors="pt").tolist()}
        out["tokenizer_image_token"][label] = cases
    json.dump(out, open(os.path.join(HERE, "prompts_ref.json"), "w"), indent=1, ensure_ascii=False, sort_keys=True)
    print("wrote prompts_ref.json:", len(out["templates"]), "templates,", 2 * len(TOKEN_CASES), "tokenisation cases")


if __name__ == "__main__":
    main()
