"""Dialogue-history retrieval (row a-R2) on the sm_100a cosine_topk kernel.

Mirrors the surface the streaming driver uses from the reference's memory_bank package:
  LocalMemoryRetrieval        memory_bank/memory_retrieval/local_doc_qa.py:180-288
    .init_cfg / .init_memory_vector_store / .load_memory_index / .search_memory
  neighbour expansion         local_doc_qa.py:135-178 (patched similarity_search_with_score_by_vector)
  JsonMemoryLoader.load       local_doc_qa.py:17-61
  enter_name / save_local_memory   memory_bank/memory_utils.py:61-109
  only_related_prompt_dict_ego     memory_bank/prompt_utils.py:37-44
  build_prompt_with_search_memory_only_related   utiles.py:1057-1078

The reference stacks langchain 0.0.142 FAISS (IndexFlatL2) over HuggingFaceEmbeddings (MiniLM-L6,
mean-pool + L2-normalise).  On unit vectors the L2 ranking equals the cosine ranking, so one
cosine_topk launch over the device-resident [n_turns, d] embedding matrix replaces the FAISS
search; the reported score is converted back to FAISS's squared L2 (2 - 2 cos).  The index is a
tensor in HBM plus the Python doc list — no on-disk FAISS files; `vs_path` keeps its role as the
cache key so the driver's rmtree/rebuild dance (memory_utils.py:76-83) still works.
"""
from __future__ import annotations

import datetime
import json
import os
import shutil
import time
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import ops

VECTOR_SEARCH_TOP_K = 3     # local_doc_qa.py:15
CHUNK_SIZE = 200            # configs/model_config.py:51

embedding_model_dict = {"minilm-l6": "sentence-transformers/all-MiniLM-L6-v2"}   # configs/model_config.py:5-18


@dataclass
class Document:
    page_content: str
    metadata: Dict[str, object] = field(default_factory=dict)


def load_memory_file(filepath: str, user_name: str, language: str = "cn") -> List[Document]:
    """One Document per dialogue turn (+ one per daily summary), local_doc_qa.py:24-61."""
    with open(filepath, "r", encoding="utf-8") as f:
        memories = json.loads(f.read())
    docs: List[Document] = []
    for name, user_memory in memories.items():
        if name != user_name:
            continue
        docs = []
        if "history" not in user_memory:
            continue
        for date, content in user_memory["history"].items():
            head = f"时间{date}的对话内容：" if language == "cn" else f"Conversation content on {date}:"
            user_kw = "[|用户|]：" if language == "cn" else "[|User|]:"
            ai_kw = "[|AI恋人|]：" if language == "cn" else "[|AI|]:"
            for dialog in content:
                text = head + f"{user_kw} {dialog['query'].strip()}; " + f"{ai_kw} {dialog['response'].strip()}"
                docs.append(Document(text, {"source": date}))
            if "summary" in user_memory and date in user_memory["summary"]:
                s = user_memory["summary"][date]
                text = f"时间{date}的对话总结为：{s}" if language == "cn" else f"The summary of the conversation on {date} is: {s}"
                docs.append(Document(text, {"source": date}))
    return docs


class MemoryIndex:
    """Device-resident flat index, APPEND-ONLY: rows live in a preallocated [capacity, d] fp32 buffer
    in HBM and new dialogue turns are written behind the live rows in place — the reference rebuilds the FAISS index
    from the whole history (and rewrites it on disk) on every question (memory_utils.py:76-83).  `embeddings` is the
    view of the live rows `cosine_topk` reads; the buffer doubles when it is full."""

    MIN_CAPACITY = 256

    def __init__(self, docs: List[Document], embeddings: torch.Tensor, embed_fn, chunk_size: int = CHUNK_SIZE):
        self.docs = list(docs)
        self.embed_fn = embed_fn
        self.chunk_size = chunk_size
        n, d = embeddings.shape
        cap = self.MIN_CAPACITY
        while cap < n:
            cap *= 2
        self._buf = torch.empty((cap, d), dtype=torch.float32, device=embeddings.device)
        self._buf[:n].copy_(embeddings)
        self.n = n
        self.reallocations = 0

    @property
    def embeddings(self) -> torch.Tensor:
        return self._buf[: self.n]

    @property
    def capacity(self) -> int:
        return self._buf.shape[0]

    def append(self, docs: Sequence[Document], rows: torch.Tensor) -> None:
        """Add documents with their unit-norm embedding rows [m, d] behind the live rows (no restack of the history)."""
        m = len(docs)
        if rows.shape[0] != m or rows.shape[1] != self._buf.shape[1]:
            raise ValueError(f"append: {m} documents but rows of shape {tuple(rows.shape)}")
        if self.n + m > self.capacity:
            cap = self.capacity
            while cap < self.n + m:
                cap *= 2
            buf = torch.empty((cap, self._buf.shape[1]), dtype=torch.float32, device=self._buf.device)
            buf[: self.n].copy_(self._buf[: self.n])
            self._buf = buf
            self.reallocations += 1
        self._buf[self.n: self.n + m].copy_(rows.to(self._buf.device, torch.float32))
        self.docs.extend(docs)
        self.n += m

    def similarity_search_with_score(self, query: str, k: int = 4) -> List[Tuple[Document, float]]:
        q = self.embed_fn([query])[0]
        scores, indices = ops.cosine_topk(self.embeddings, q.to(self.embeddings.device), k=k)
        scores, indices = scores[0].tolist(), indices[0].tolist()          # one D2H read per question
        n = len(self.docs)
        out: List[Tuple[Document, float]] = []
        id_set = set()
        last_j = 0
        for j, i in enumerate(indices):
            if i == -1:
                continue
            last_j = j
            doc = self.docs[i]
            id_set.add(i)
            docs_len = len(doc.page_content)
            # neighbour expansion, local_doc_qa.py:149-163 (`len(docs)` there is the still-empty
            # OUTPUT list, so the range is 1..i-1; `break` leaves only the inner two-element loop)
            for kk in range(1, max(i, len(out) - i)):
                for l in (i + kk, i - kk):
                    if 0 <= l < n:
                        doc0 = self.docs[l]
                        if docs_len + len(doc0.page_content) > self.chunk_size:
                            break
                        elif doc0.metadata["source"] == doc.metadata["source"]:
                            docs_len += len(doc0.page_content)
                            id_set.add(l)
        if not id_set:
            return out
        runs, run = [], []
        for idx in sorted(id_set):
            if run and idx != run[-1] + 1:
                runs.append(run)
                run = []
            run.append(idx)
        runs.append(run)
        l2 = 2.0 - 2.0 * scores[last_j]                                    # FAISS IndexFlatL2 reports squared L2
        for seq in runs:
            # the reference concatenates into the stored document in place (:170); a copy is returned
            # here so repeated searches on one index do not accumulate text
            merged = Document("".join(self.docs[i].page_content for i in seq), dict(self.docs[seq[0]].metadata))
            out.append((merged, l2))
        return out


class LocalMemoryRetrieval:
    embeddings = None
    top_k: int = VECTOR_SEARCH_TOP_K
    chunk_size: int = CHUNK_SIZE
    _stores: Dict[str, MemoryIndex] = {}

    def init_cfg(self, embedding_model: str = "minilm-l6", embedding_device="cuda", top_k=VECTOR_SEARCH_TOP_K,
                 language="cn", embed_fn: Optional[Callable[[Sequence[str]], torch.Tensor]] = None):
        """`embed_fn(list[str]) -> [n, d]` unit-norm embeddings on the device.  Without it the sentence encoder named
        by `embedding_model` (a key of `embedding_model_dict` or a local checkpoint directory) is loaded onto the
        sm_100a kernels: BERT encoder -> mean pool -> L2 normalise, what langchain's HuggingFaceEmbeddings computes."""
        self.language = language
        self.top_k = top_k
        self.device = torch.device(embedding_device if str(embedding_device).startswith("cuda") else "cuda:0")
        if embed_fn is None:
            embed_fn = _hf_sentence_embedder(embedding_model_dict.get(embedding_model, embedding_model), self.device)
        self.embeddings = embed_fn

    def init_memory_vector_store(self, filepath, vs_path=None, user_name=None, cur_date=None):
        files = [filepath] if isinstance(filepath, str) else list(filepath)
        if isinstance(filepath, str) and not os.path.exists(filepath):
            return None, None
        if isinstance(filepath, str) and os.path.isdir(filepath):
            files = [os.path.join(filepath, f) for f in os.listdir(filepath)]
        docs, loaded = [], []
        for f in files:
            try:
                docs += load_memory_file(f, user_name, self.language)
                loaded.append(f)
            except Exception as e:  # noqa: BLE001  (the reference swallows per-file errors too, :226-229)
                print(e)
        if not docs:
            return None, loaded
        if not vs_path:
            vs_path = f"memory_index_{datetime.datetime.now().strftime('%Y%m%d_%H%M%S')}"
        prev = self._stores.get(vs_path)
        texts = [d.page_content for d in docs]
        if prev is not None and os.path.isdir(vs_path):                     # FAISS.load_local + add_documents (:238-241)
            prev.append(docs, self._embed_cached(texts))
        elif prev is not None and len(prev.docs) <= len(docs) and \
                all(a.page_content == b for a, b in zip(prev.docs, texts)) and \
                all(a.metadata == b.metadata for a, b in zip(prev.docs, docs)):
            # the driver deletes the index directory and rebuilds it from the whole history before every question
            # (memory_utils.py:76-83): the history only ever GROWS, so the resident index is kept and the new turns
            # are appended in place — same content as the rebuild, O(new turns) work
            tail = docs[len(prev.docs):]
            if tail:
                prev.append(tail, self._embed_cached(texts[len(prev.docs):]))
        else:
            self._stores[vs_path] = MemoryIndex(docs, self._embed_cached(texts), self.embeddings, self.chunk_size)
        os.makedirs(vs_path, exist_ok=True)                                 # keeps the driver's exists()/rmtree logic valid
        return vs_path, loaded

    # ---- O(new turns) per question instead of O(history) -------------------------------------
    # The reference re-embeds the whole dialogue history and rewrites the FAISS directory on every question
    # (memory_utils.py:76-83).  Same results here with (1) an embedding cache keyed by the turn text, so only
    # turns not seen before go through the encoder, and (2) an explicit snapshot of the device-resident index
    # (one file, written when the caller asks) instead of a per-question rewrite.
    _embed_cache: Dict[str, torch.Tensor] = {}
    embed_calls: int = 0       # texts actually sent to the encoder (tests / accounting)

    def _embed_cached(self, texts: Sequence[str]) -> torch.Tensor:
        cache = self._embed_cache
        new = [t for t in dict.fromkeys(texts) if t not in cache]
        if new:
            rows = self.embeddings(new).to(self.device, torch.float32)
            self.embed_calls += len(new)
            for t, r in zip(new, rows):
                cache[t] = r
        return torch.stack([cache[t] for t in texts]).contiguous()

    def save_snapshot(self, vs_path: str) -> str:
        """Write the index held for `vs_path` (documents + fp32 embeddings) to `<vs_path>/index.pt`."""
        store = self._stores[vs_path]
        os.makedirs(vs_path, exist_ok=True)
        path = os.path.join(vs_path, "index.pt")
        torch.save({"docs": [(d.page_content, d.metadata) for d in store.docs],
                    "embeddings": store.embeddings.detach().cpu(), "chunk_size": store.chunk_size}, path)
        return path

    def load_memory_index(self, vs_path):
        store = self._stores.get(vs_path)
        if store is None:                                                   # new process: resume from the snapshot
            path = os.path.join(vs_path, "index.pt")
            if not os.path.exists(path):
                raise KeyError(f"no memory index for {vs_path!r} (neither in this process nor snapshotted)")
            blob = torch.load(path, map_location="cpu")
            docs = [Document(t, dict(m)) for t, m in blob["docs"]]
            store = MemoryIndex(docs, blob["embeddings"].to(self.device, torch.float32).contiguous(), self.embeddings,
                                blob.get("chunk_size", self.chunk_size))
            self._stores[vs_path] = store
        store.chunk_size = self.chunk_size
        return store

    def search_memory(self, query, vector_store):
        related = vector_store.similarity_search_with_score(query, k=self.top_k)
        docs = []
        for doc, score in related:
            doc.metadata["score"] = score
            docs.append(doc)
        docs = sorted(docs, key=lambda d: d.metadata["source"])
        pre_date, date_docs, dates = "", [], []
        for doc in docs:
            doc.page_content = doc.page_content.replace(f'时间{doc.metadata["source"]}的对话内容：', "").strip()
            if doc.metadata["source"] != pre_date:
                date_docs.append(doc.page_content)
                pre_date = doc.metadata["source"]
                dates.append(pre_date)
            else:
                date_docs[-1] += f"\n{doc.page_content}"
        return date_docs, ", ".join(dates)


def _hf_sentence_embedder(model_name: str, device: torch.device):
    """The dialogue-memory encoder (all-MiniLM-L6-v2, local_doc_qa.py:185-194) on the sm_100a kernels
    (text_encoder.SentenceEncoder).  `model_name` is a local checkpoint directory, or a hub id that
    `huggingface_hub.snapshot_download` can resolve from the local cache (there is no network at run time)."""
    import os

    from .text_encoder import SentenceEncoder
    path = model_name
    if not os.path.isdir(path):
        from huggingface_hub import snapshot_download
        path = snapshot_download(model_name, local_files_only=True)
    return SentenceEncoder.from_pretrained(path, device)


# ---- memory_bank/memory_utils.py ----------------------------------------------------------------
def enter_name(name, memory, local_memory_qa, data_args, update_memory_index=True):
    """memory_utils.py:61-90 (gradio State unwrapping dropped: the driver passes plain objects)."""
    cur_date = datetime.date.today().strftime("%Y-%m-%d")
    user_memory_index = None
    memory_dir = os.path.join(data_args.memory_basic_dir, data_args.memory_file)
    if name in memory.keys():
        user_memory = memory[name]
        memory_index_path = os.path.join(data_args.memory_basic_dir, f"memory_index/{name}_index")
        os.makedirs(os.path.dirname(memory_index_path), exist_ok=True)
        if (not os.path.exists(memory_index_path)) or update_memory_index:
            if os.path.exists(memory_index_path):
                shutil.rmtree(memory_index_path)
            memory_index_path, _ = local_memory_qa.init_memory_vector_store(
                filepath=memory_dir, vs_path=memory_index_path, user_name=name, cur_date=cur_date)
        user_memory_index = local_memory_qa.load_memory_index(memory_index_path) if memory_index_path else None
        msg = f"欢迎回来，{name}！" if data_args.language == "cn" else f"Wellcome Back, {name}！"
        return msg, user_memory, memory, name, user_memory_index
    memory[name] = {}
    memory[name].update({"name": name})
    msg = (f"欢迎新用户{name}！我会记住你的名字，下次见面就能叫你的名字啦！" if data_args.language == "cn"
           else f"Welcome, new user {name}! I will remember your name, so next time we meet, I'll be able to call you by your name!")
    return msg, memory[name], memory, name, user_memory_index


def save_local_memory(memory, b, user_name, data_args):
    """memory_utils.py:95-109."""
    memory_dir = os.path.join(data_args.memory_basic_dir, data_args.memory_file)
    date = time.strftime("%Y-%m-%d", time.localtime())
    if memory[user_name].get("history") is None:
        memory[user_name].update({"history": {}})
    if memory[user_name]["history"].get(date) is None:
        memory[user_name]["history"][date] = []
    memory[user_name]["history"][date].append({"query": b[-1][0], "response": b[-1][1]})
    json.dump(memory, open(memory_dir, "w", encoding="utf-8"), ensure_ascii=False, indent=4)
    return memory


def only_related_prompt_dict_ego():
    """memory_bank/prompt_utils.py:37-44."""
    return {"en": """
    Based on the current user's question, the most relevant historical contextual conversation records are: "{related_memory_content}".
    """}


def build_prompt_with_search_memory_only_related(text, user_name, user_memory_index, local_memory_qa, meta_prompt,
                                                 user_keyword, ai_keyword, boot_actual_name):
    """utiles.py:1057-1078."""
    query = text.replace(user_keyword, user_name).replace(ai_keyword, "AI")
    if user_memory_index:
        related_memos, _ = local_memory_qa.search_memory(query, user_memory_index)
        related_memory_content = f"\n{str(chr(10).join(related_memos)).strip()}\n"
        return meta_prompt.format(related_memory_content=related_memory_content)
    return None
