"""streamchat_b200 — B200-native (sm_100a) implementation of StreamChat's per-frame hot path.

Only the hot path named in BASELINE.json lives here: CLIP-ViT frame encode, k-means selective
stacking, cosine top-k memory retrieval and LongVA (Qwen2) prefill/decode, behind the
reference's own Python call surface (see SURVEY.md §8 and DESIGN.md).
"""
__version__ = "0.1.0"
