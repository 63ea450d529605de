"""ChatML conversation templates used on the hot path (reference longva/conversation.py:
Conversation :24-110 CHATML branch :86-96; templates conv_qwen :414-423, conv_qwen_ego :525-538,
conv_qwen_captioning :540-550, conv_qwen_summarize :553-562; registry :587-590).
Only the qwen / ChatML family is provided — the other separator styles belong to model families
that are out of scope."""
from __future__ import annotations

import copy as _copy
from dataclasses import dataclass, field
from enum import Enum, auto
from typing import List, Optional, Sequence


class SeparatorStyle(Enum):
    CHATML = auto()


@dataclass
class Conversation:
    system: str
    roles: Sequence[str]
    messages: List[List[Optional[str]]] = field(default_factory=list)
    offset: int = 0
    sep_style: SeparatorStyle = SeparatorStyle.CHATML
    sep: str = "<|im_end|>"
    sep2: Optional[str] = None
    version: str = "qwen"

    def append_message(self, role: str, message: Optional[str]) -> None:
        self.messages.append([role, message])

    def get_prompt(self) -> str:
        if self.sep_style != SeparatorStyle.CHATML:
            raise NotImplementedError("only the ChatML separator style is implemented")
        text = "" if self.system == "" else self.system + self.sep + "\n"
        for role, message in self.messages:
            if message:
                if isinstance(message, tuple):            # (text, images) form
                    message = "<image>" * len(message[1]) + message[0]
                text += role + "\n" + message + self.sep + "\n"
            else:
                text += role + "\n"                       # open assistant turn
        return text

    def copy(self) -> "Conversation":
        return Conversation(system=self.system, roles=self.roles,
                            messages=[[r, m] for r, m in self.messages], offset=self.offset,
                            sep_style=self.sep_style, sep=self.sep, sep2=self.sep2, version=self.version)


_ROLES = ("<|im_start|>user", "<|im_start|>assistant")

conv_qwen = Conversation(system="<|im_start|>system\nYou are a helpful assistant.", roles=_ROLES)

conv_qwen_ego = Conversation(
    system="<|im_start|>system\n    You are a useful assistant. What you see is video from my first-person "
           "perspective and you need to conduct multiple rounds of dialogue with me.", roles=_ROLES)

conv_qwen_captioning = Conversation(
    system="<|im_start|>system\n    You are a useful visual assistant. Please describe what you see in this "
           "video in as much detail as possible from a first-person perspective, including the surrounding "
           "environment, what objects are there, etc.\n    PLEASE DO NOT GENERATE TEXT YOU ARE NOT SURE ABOUT.",
    roles=_ROLES)

conv_qwen_summarize = Conversation(system="<|im_start|>system\n    You are a helpful assistant.", roles=_ROLES)

conv_templates = {
    "qwen_1_5": conv_qwen,
    "qwen_1_5_ego": conv_qwen_ego,
    "qwen_1_5_caption": conv_qwen_captioning,
    "qwen_1_5_summarize": conv_qwen_summarize,
}
default_conversation = conv_qwen
