"""CPU oracle for dialogue retrieval (TEST INFRASTRUCTURE — see oracle/__init__.py).

Restates the stack the reference reaches at memory_bank/memory_retrieval/local_doc_qa.py:263-288:
langchain==0.0.142 `FAISS.similarity_search_with_score` (requirements.txt:63; embeds the query,
calls the patched `similarity_search_with_score_by_vector`, local_doc_qa.py:135-178) over
faiss-cpu==1.8.0.post1 `IndexFlatL2.search` (requirements.txt:32: exact brute force, squared L2,
ascending).  None of the three packages is vendored or installed: PARITY UNPINNED beyond the
reference's own call sites; embeddings are injected.
"""
from __future__ import annotations

import copy
from typing import List

import numpy as np


def index_flat_l2_search(xb: np.ndarray, xq: np.ndarray, k: int):
    """faiss.IndexFlatL2.search: squared L2 distances, k smallest, -1 padding."""
    d = ((xq[:, None, :].astype(np.float32) - xb[None, :, :].astype(np.float32)) ** 2).sum(-1)
    order = np.argsort(d, axis=1, kind="stable")[:, :k]
    scores = np.take_along_axis(d, order, axis=1)
    if xb.shape[0] < k:
        pad = k - xb.shape[0]
        order = np.concatenate([order, -np.ones((xq.shape[0], pad), dtype=order.dtype)], 1)
        scores = np.concatenate([scores, np.full((xq.shape[0], pad), np.inf, dtype=np.float32)], 1)
    return scores, order


def similarity_search_with_score_by_vector(docs_store: List, xb: np.ndarray, embedding: np.ndarray, k: int,
                                           chunk_size: int):
    """local_doc_qa.py:135-178 on a list of objects with .page_content / .metadata['source']."""
    scores, indices = index_flat_l2_search(xb, embedding[None, :], k)
    docs = []
    id_set = set()
    j = 0
    for j, i in enumerate(indices[0]):
        if i == -1:
            continue
        doc = docs_store[i]
        id_set.add(int(i))
        docs_len = len(doc.page_content)
        for kk in range(1, max(int(i), len(docs) - int(i))):
            for l in [int(i) + kk, int(i) - kk]:
                if 0 <= l < len(docs_store):
                    doc0 = docs_store[l]
                    if docs_len + len(doc0.page_content) > chunk_size:
                        break
                    elif doc0.metadata["source"] == doc.metadata["source"]:
                        docs_len += len(doc0.page_content)
                        id_set.add(l)
    id_list = sorted(id_set)
    runs, run = [], [id_list[0]]
    for a, b in zip(id_list, id_list[1:]):
        if a + 1 == b:
            run.append(b)
        else:
            runs.append(run)
            run = [b]
    runs.append(run)
    for seq in runs:
        doc = copy.deepcopy(docs_store[seq[0]])
        for idx in seq[1:]:
            doc.page_content += docs_store[idx].page_content
        docs.append((doc, float(scores[0][j])))
    return docs


def search_memory(docs_store, xb, query_embedding, top_k, chunk_size=200):
    """local_doc_qa.py:263-288."""
    related = similarity_search_with_score_by_vector(docs_store, xb, query_embedding, top_k, chunk_size)
    docs = []
    for doc, score in related:
        doc.metadata["score"] = score
        docs.append(doc)
    docs = sorted(docs, key=lambda d: d.metadata["source"])
    pre, date_docs, dates = "", [], []
    for doc in docs:
        doc.page_content = doc.page_content.replace(f'时间{doc.metadata["source"]}的对话内容：', "").strip()
        if doc.metadata["source"] != pre:
            date_docs.append(doc.page_content)
            pre = doc.metadata["source"]
            dates.append(pre)
        else:
            date_docs[-1] += f"\n{doc.page_content}"
    return date_docs, ", ".join(dates)
