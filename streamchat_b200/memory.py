"""Host-side mirrors of the reference's memory functions (utiles.py), same names and argument
order, with the arithmetic on the sm_100a kernels.

  calculate_forgetting_probabilities      <- utiles.py:251-255  (a-S1)
  select_data_without_replacement         <- utiles.py:257-262  (a-S2)
  MultimodalTreeNode                      <- utiles.py:48-56    (a-T1)
  count_nodes_by_depth                    <- utiles.py:1002-1011
  fast_building_memory_tree_summarize_token <- utiles.py:489-620 (a-T2)
  fast_search_tree_multi_modal_with_embedding <- utiles.py:685-788 (a-R1)

Differences from the reference, all deliberate and listed in DESIGN.md:
  * node text embeddings are computed once per node and cached on the node (`text_embedding`)
    instead of one batch-1 BERT forward + D2H sync per visited child per question; the per-level
    similarity + arg-max is one `cosine_topk` launch over the stacked child embeddings;
  * captions can be generated greedily (`caption_do_sample=False`) for parity runs (hazard H2).
"""
from __future__ import annotations

from collections import defaultdict
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch

from . import ops
from .kmeans import weighted_kmeans_feature

RED = "\033[31m"
BLUE = "\033[34m"
RESET = "\033[0m"

VERBOSE = False  # the reference prints the whole tree on every update (utiles.py:491-504)
CAPTION_TOKENS = 128   # max_new_tokens of the per-chunk caption (utiles.py:555)
SUMMARY_TOKENS = 256   # max_new_tokens of the merge summary (utiles.py:604)
# The reference samples captions / summaries at T=0.1 (utiles.py:551-552,600-601).  The driver calls
# fast_building_memory_tree_summarize_token positionally (inference_streaming_longva_v2.py:351), so the greedy
# override of parity runs (hazard H2) is a module switch as well as a keyword.
CAPTION_DO_SAMPLE = True


class MultimodalTreeNode:
    """utiles.py:48-56.  `text` may be PENDING: captions and merge summaries are greedy generations
    whose result nothing reads until the next summary prompt or a tree search, and neither the tree
    shape nor the centroids depend on them — so they are recorded as recipes and generated later in
    large lock-step batches (`resolve_pending`; same ids, one weight pass per decode step for up to
    `max_batch` sequences).  Reading `.text` earlier resolves the node (and what it depends on)."""

    def __init__(self, centroids, text, text_distance=None, image_distance=None, labels=None, depth=0):
        self.centroids = centroids
        self._text = text
        self._pending = None        # dict recipe, see fast_building_memory_tree_summarize_token
        self.text_distance = text_distance
        self.image_distance = image_distance
        self.labels = labels
        self.children = []
        self.depth = depth
        self.text_embedding = None  # cached [D] embedding of `text` (not in the reference)

    @property
    def text(self):
        if self._pending is not None:
            resolve_pending([self])
        return self._text

    @text.setter
    def text(self, value):
        self.set_text(value)

    def set_text(self, value):
        self._text = value
        self._pending = None
        self.text_embedding = None


def _collect_pending(nodes, out):
    for n in nodes:
        if n._pending is not None:
            out.append(n)
        _collect_pending(n.children, out)


def resolve_pending(nodes) -> int:
    """Generate every pending caption / summary reachable from `nodes`: captions first (batched),
    then summaries level by level (a summary prompt quotes its children's texts, utiles.py:506-523).
    Returns the number of texts generated."""
    todo: list = []
    _collect_pending(nodes, todo)
    todo = list({id(n): n for n in todo}.values())
    done = 0
    while todo:
        ready = [n for n in todo if n._pending["kind"] == "caption"]
        if not ready:
            ready = [n for n in todo if all(c._pending is None for c in n._pending["children"])]
        if not ready:
            raise RuntimeError("pending memory texts form a cycle")
        summarizer, tokenizer = ready[0]._pending["summarizer"], ready[0]._pending["tokenizer"]
        reqs = []
        for n in ready:
            r = n._pending
            if r["kind"] == "caption":
                reqs.append((r["input_ids"].to(summarizer.device), [r["features"]], r["max_new"]))
            else:
                ids = make_summary_prompt([c._text for c in r["children"]], tokenizer, r["conv_templates"])
                reqs.append((ids.to(summarizer.device), None, r["max_new"]))
        with torch.no_grad():
            ids_list = summarizer.generate_requests(reqs, modalities=["video"])
        for n, ids in zip(ready, ids_list):
            n.set_text(tokenizer.batch_decode(ids, skip_special_tokens=True)[0].strip())
        done += len(ready)
        ready_ids = {id(n) for n in ready}
        todo = [n for n in todo if id(n) not in ready_ids]
    return done


def texts_read_by_search(all_nodes) -> list:
    """The nodes whose `.text` a tree search may read (utiles.py:709-764): every depth-0 root and everything BELOW a
    deeper root — never the deeper roots' own summaries, which therefore may stay pending across a question."""
    out = []
    for n in all_nodes:
        if n.depth == 0:
            out.append(n)
        else:
            out.extend(n.children)
    return out


def _all_nodes(nodes) -> list:
    out = []
    for n in nodes:
        out.append(n)
        out.extend(_all_nodes(n.children))
    return out


def ready_summaries(nodes) -> list:
    """Pending merge summaries (anywhere under `nodes`) whose children's texts all exist: they can be generated now,
    e.g. in the same lock-step batch as an answer (stream.StreamingSession.answer)."""
    todo: list = []
    _collect_pending(nodes, todo)
    todo = list({id(n): n for n in todo}.values())
    return [n for n in todo if n._pending["kind"] == "summary" and all(c._pending is None for c in n._pending["children"])]


def summary_request(node):
    """`(input_ids, None, max_new_tokens)` of a ready pending summary (utiles.py:506-523,594-605)."""
    r = node._pending
    ids = make_summary_prompt([c._text for c in r["children"]], r["tokenizer"], r["conv_templates"])
    return ids, None, r["max_new"]


def calculate_forgetting_probabilities(length, tau=10):
    t = np.arange(length)
    R_t = np.exp(-t / tau)
    return R_t / R_t.sum()


def select_data_without_replacement(queue, probabilities, selection_length=10):
    indices = np.arange(len(queue))
    selected_indices = np.random.choice(indices, size=selection_length, replace=False, p=probabilities)
    return [queue[idx] for idx in selected_indices]


def count_nodes_by_depth(nodes):
    depth_count = defaultdict(int)
    for node in nodes:
        depth_count[node.depth] += 1
        stack = node.children.copy()
        while stack:
            current_node = stack.pop()
            depth_count[current_node.depth] += 1
            stack.extend(current_node.children)
    return depth_count


def _print_tree(node, indent=0):
    if isinstance(node, list):
        for n in node:
            _print_tree(n, indent)
        return
    print(" " * indent + f"{RED}Depth{RESET}: {node.depth}, {RED}Centroids{RESET}: "
          f"{tuple(node.centroids.shape)}, {RED}Text{RESET} :{node.text}")
    for child in node.children:
        _print_tree(child, indent + 4)


def get_summarize_depth(nodes, interval):
    """utiles.py:525-536: highest depth whose node count is a positive multiple of `interval`."""
    depth_count = defaultdict(int)
    for node in nodes:
        depth_count[node.depth] += 1
    max_depth = max(depth_count.keys())
    for depth in range(max_depth, -1, -1):
        if depth_count[depth] % interval == 0 and depth_count[depth] > 0:
            return depth, depth_count
    return 0, depth_count


def make_summary_prompt(caption_list, tokenizer, conv_templates):
    """utiles.py:506-523."""
    order = ["first", "second", "third", "fourth", "fifth", "sixth", "seventh", "eighth", "ninth", "tenth"]
    new_caption = ["The caption of the {} video clip is:{} \n".format(order[i], c) for i, c in enumerate(caption_list)]
    qs = " ".join(new_caption)
    qs = ("You need to write a summary of the following, including as many key details as possible "
          "into one sentence." + qs)
    conv = conv_templates["qwen_1_5_summarize"].copy()
    conv.append_message(conv.roles[0], qs)
    conv.append_message(conv.roles[1], None)
    summarize_prompt = conv.get_prompt()
    return torch.tensor(tokenizer(summarize_prompt).input_ids, dtype=torch.long).unsqueeze(0)


def fast_building_memory_tree_summarize_token(k_means_chunk_feature_list, num_clusters, interval, summarizer,
                                              input_ids, tokenizer, chunked_feature_list, existing_tree=None,
                                              *, caption_do_sample=None, kmeans_init_indices=None,
                                              conv_templates=None, lazy_text=True, chunk_nodes=None):
    """One memory-tree update: caption every new chunk, append depth-0 nodes, perform at most one
    merge (k-means over the merged nodes' frames + text summary).  Same positional signature as
    utiles.py:489; keyword-only extras inject determinism.  With a greedy captioner that offers
    `generate_requests` (the sm_100a engine) and `lazy_text`, captions and summaries are deferred and
    generated in lock-step batches by `resolve_pending` — identical ids, far fewer weight passes.
    `chunk_nodes`: depth-0 nodes built elsewhere for these chunks (the frame-sharded stream captions a chunk on the
    rank that owns it and gathers the result, distributed.ShardedStreamingSession) — used instead of captioning."""
    if conv_templates is None:
        from .longva.conversation import conv_templates as _ct
        conv_templates = _ct
    if caption_do_sample is None:
        caption_do_sample = CAPTION_DO_SAMPLE
    flat_chunks = []
    for chunk_feature in (chunked_feature_list if chunk_nodes is None else []):
        dimension = chunk_feature[0].shape[-1]
        flat_chunks.append(torch.cat(chunk_feature, dim=0).reshape(-1, dimension).to(summarizer.device))
    lazy = (not caption_do_sample) and hasattr(summarizer, "generate_requests") and lazy_text
    nodes = []
    if chunk_nodes is not None:
        if len(chunk_nodes) != len(k_means_chunk_feature_list):
            raise ValueError("chunk_nodes must hold one depth-0 node per chunk")
        nodes = list(chunk_nodes)
    elif lazy:
        # greedy captions are recipes until somebody needs the text (see MultimodalTreeNode)
        for t, feats in zip(k_means_chunk_feature_list, flat_chunks):
            n = MultimodalTreeNode(t, None, depth=0)
            n._pending = {"kind": "caption", "summarizer": summarizer, "tokenizer": tokenizer,
                          "input_ids": input_ids, "features": feats, "max_new": CAPTION_TOKENS}
            nodes.append(n)
    else:
        for t, chunk_feature in zip(k_means_chunk_feature_list, flat_chunks):
            with torch.no_grad():
                output_ids = summarizer.generate_with_image_embedding(
                    input_ids.to(summarizer.device), image_embeddings=[chunk_feature], modalities=["video"],
                    do_sample=caption_do_sample, temperature=0.1 if caption_do_sample else 0, top_p=None,
                    max_new_tokens=CAPTION_TOKENS, use_cache=False)
            nodes.append(MultimodalTreeNode(t, tokenizer.batch_decode(output_ids, skip_special_tokens=True)[0].strip(),
                                            depth=0))
    if existing_tree:
        nodes = existing_tree + nodes

    summarize_depth, _ = get_summarize_depth(nodes, interval)
    start_index = next((i for i, n in enumerate(nodes) if n.depth == summarize_depth), None)
    chunk_length = len([x for x in nodes if x.depth == summarize_depth])
    if VERBOSE:
        print("summarize_depth:{}/ start_index:{}/ chunk_length:{} / len(nodes):{}".format(
            summarize_depth, start_index, chunk_length, len(nodes)))

    if chunk_length % interval >= 0 and len(nodes) > 0 and chunk_length >= interval:   # utiles.py:575
        chunk = nodes[start_index: start_index + interval]
        combined_centroids = torch.cat([n.centroids for n in chunk], dim=0)
        if combined_centroids.shape[0] > num_clusters:
            new_centroids, _ = weighted_kmeans_feature(combined_centroids, num_clusters,
                                                       init_indices=kmeans_init_indices)
        else:
            new_centroids = combined_centroids
        if lazy:
            new_node = MultimodalTreeNode(new_centroids, None, depth=chunk[0].depth + 1)
            new_node._pending = {"kind": "summary", "summarizer": summarizer, "tokenizer": tokenizer,
                                 "children": list(chunk), "conv_templates": conv_templates,
                                 "max_new": SUMMARY_TOKENS}
        else:
            summarize_ids = make_summary_prompt([n.text for n in chunk], tokenizer, conv_templates)
            with torch.no_grad():
                output_ids = summarizer.generate_with_image_embedding(
                    summarize_ids.to(summarizer.device), image_embeddings=None, modalities=["video"],
                    do_sample=caption_do_sample, temperature=0.1 if caption_do_sample else 0, top_p=None,
                    max_new_tokens=SUMMARY_TOKENS, use_cache=False)
            summarize_text = tokenizer.batch_decode(output_ids, skip_special_tokens=True)[0].strip()
            new_node = MultimodalTreeNode(new_centroids, summarize_text, depth=chunk[0].depth + 1)
        for node in chunk:
            new_node.children.append(node)
        nodes[start_index: start_index + interval] = [new_node]
    if VERBOSE:
        _print_tree(nodes)
    return nodes


# ------------------------------------------------------------------------------------------
def embed_text(model, tokenizer, text: str) -> torch.Tensor:
    """CLS embedding of `text` (utiles.py:704-708: `model(**ids).last_hidden_state[:, 0]`).
    `model` may also be any callable `(tokenizer, text) -> Tensor[D]` (used by tests / synthetic
    streams where no BERT weights exist)."""
    if not hasattr(model, "forward") and callable(model):
        return model(tokenizer, text)
    ids = tokenizer(text, padding=True, return_tensors="pt")
    dev = next(model.parameters()).device
    ids = {k: v.to(dev) for k, v in ids.items()}
    with torch.no_grad():
        return model(**ids).last_hidden_state[:, 0][0]


def _node_embedding(node, model, tokenizer) -> torch.Tensor:
    if node.text_embedding is None:
        node.text_embedding = embed_text(model, tokenizer, node.text).detach()
    return node.text_embedding


def node_embeddings(nodes: Sequence[MultimodalTreeNode], model, tokenizer) -> torch.Tensor:
    """[len(nodes), D] text embeddings of `nodes`, cached on the nodes.  Pending texts are generated first
    in lock-step batches (never one node at a time), then every uncached text goes through ONE packed
    forward of the on-device encoder (text_encoder.BertTextEncoder.encode_texts)."""
    resolve_pending(list(nodes))
    missing = [n for n in nodes if n.text_embedding is None]
    if len(missing) > 1 and hasattr(model, "encode_texts"):
        for n, e in zip(missing, model.encode_texts(tokenizer, [n.text for n in missing])):
            n.text_embedding = e
    return torch.stack([_node_embedding(n, model, tokenizer) for n in nodes])


def _best_by_cosine(nodes: Sequence[MultimodalTreeNode], query_embedding: torch.Tensor, model, tokenizer,
                    default):
    """Reference rule (utiles.py:719-742): best_sim starts at 0, strict '>' — first maximum wins,
    non-positive similarities never win.  One cosine_topk launch replaces the Python loop."""
    index = node_embeddings(nodes, model, tokenizer).to(query_embedding.device)
    scores, idx = ops.cosine_topk(index, query_embedding, k=1)
    s, i = float(scores[0, 0]), int(idx[0, 0])
    if VERBOSE:
        print(f"{RED}sim{RESET}:{s} best child {i}")
    return i if s > 0 else default


def fast_search_tree_multi_modal_with_embedding(all_nodes, query, image_embedding, model, tokenizer, top_k=1):
    """Walk every non-leaf root down its most similar child chain, then pick the most similar
    depth-0 root.  `image_embedding` and `top_k` are accepted and ignored, as in the reference."""
    path_features: List[torch.Tensor] = []
    path_text: List[str] = []
    redundant_nodes = []
    resolve_pending(texts_read_by_search(all_nodes))   # every text the walk may read, in batched generations
    query_embedding = embed_text(model, tokenizer, query)
    for node in all_nodes:
        current_node = node
        if current_node.depth == 0:
            redundant_nodes.append(node)
        else:
            while current_node.children:
                best_child_index = _best_by_cosine(current_node.children, query_embedding, model, tokenizer, None)
                # the reference indexes children[None] here and raises TypeError when no child has a
                # positive similarity (utiles.py:746); list indexing with None reproduces that.
                path_features.append(current_node.children[best_child_index].centroids)
                path_text.append(current_node.children[best_child_index].text)
                current_node = current_node.children[best_child_index]
    if len(redundant_nodes) >= 1:
        best_index = _best_by_cosine(redundant_nodes, query_embedding, model, tokenizer, 0)   # utiles.py:751
        path_features.append(redundant_nodes[best_index].centroids)
        path_text.append(redundant_nodes[best_index].text)
    return path_features, path_text
