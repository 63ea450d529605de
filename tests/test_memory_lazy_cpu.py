"""CPU: the lazy memory texts of streamchat_b200/memory.py (host logic only, fake captioner).

Captions (utiles.py:545-556) and merge summaries (:594-605) are greedy generations nobody reads until the
next summary prompt or a tree search, so the tree update records them as recipes and `resolve_pending`
generates them in lock-step batches: all captions of the tree in one batch, then the summaries level by
level (a summary prompt quotes its children's texts, :506-523).  Tree shape and centroids never depend on them."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fakes  # noqa: E402

from streamchat_b200 import memory as M  # noqa: E402


class BatchSummarizer(fakes.FakeSummarizer):
    """Offers the engine's batched entry point; records what each batch contained."""

    def __init__(self):
        super().__init__("cpu")
        self.batches = []

    def generate_requests(self, reqs, modalities=None):
        self.batches.append(["caption" if feats is not None else "summary" for _, feats, _ in reqs])
        outs = []
        for ids, feats, max_new in reqs:
            outs.append(self.generate_with_image_embedding(ids, image_embeddings=feats, max_new_tokens=max_new))
        return outs


def _update(tree, summ, tok, n_chunks, interval):
    feats = [[torch.randn(1, 4, 8) for _ in range(3)] for _ in range(n_chunks)]          # 3 frames per chunk
    kmeans = [torch.cat(ch) for ch in feats]
    ids = torch.tensor([[5, 6, -200, 7]])
    return M.fast_building_memory_tree_summarize_token(kmeans, 1000, interval, summ, ids, tok, feats, tree,
                                                       caption_do_sample=False, lazy_text=True)


def test_updates_generate_nothing_and_resolve_batches_by_level():
    torch.manual_seed(0)
    summ, tok = BatchSummarizer(), fakes.FakeTokenizer()
    tree = None
    for _ in range(4):                       # 4 updates of 1 chunk, merge every 2 nodes (num_clusters large: no k-means)
        tree = _update(tree, summ, tok, 1, 2)
    # at most one merge per update (utiles.py:575): [c0], [m(c0,c1)], [m, c2], [m, c2, c3] -> m(c2,c3) -> ...
    assert summ.calls == 0 and summ.batches == []                       # nothing was generated yet
    depths = sorted(n.depth for n in tree)
    n_done = M.resolve_pending(tree)
    kinds = summ.batches
    assert kinds[0] == ["caption"] * 4                                   # every caption of the tree in ONE batch
    assert all(set(b) == {"summary"} for b in kinds[1:]) and len(kinds) >= 2
    assert n_done == 4 + sum(len(b) for b in kinds[1:])
    assert M.resolve_pending(tree) == 0                                  # idempotent
    assert all(n._pending is None and isinstance(n.text, str) and n.text for n in tree)
    assert depths == sorted(n.depth for n in tree)                       # resolving texts does not touch the shape


def test_reading_one_text_resolves_only_what_it_needs():
    torch.manual_seed(1)
    summ, tok = BatchSummarizer(), fakes.FakeTokenizer()
    tree = _update(None, summ, tok, 3, 10)                               # three depth-0 nodes, no merge
    _ = tree[1].text
    assert summ.batches == [["caption"]] and tree[0]._pending is not None and tree[2]._pending is not None
    tree[0].text = "given"                                               # assigning a text cancels the recipe
    assert tree[0]._pending is None and tree[0].text == "given" and summ.batches == [["caption"]]


def test_eager_path_without_batched_entry_point_matches_lazy_texts():
    """A captioner without `generate_requests` (the reference's HF model) runs the reference's eager order."""
    torch.manual_seed(2)
    tok = fakes.FakeTokenizer()
    eager = fakes.FakeSummarizer("cpu")
    torch.manual_seed(3)
    t_eager = _update(None, eager, tok, 3, 10)
    assert eager.calls == 3 and all(n._pending is None for n in t_eager)
    lazy = BatchSummarizer()
    torch.manual_seed(3)
    t_lazy = _update(None, lazy, tok, 3, 10)
    M.resolve_pending(t_lazy)
    assert [n.text for n in t_lazy] == [n.text for n in t_eager]         # same call order -> same fake ids
