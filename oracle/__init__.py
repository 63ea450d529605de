"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatements (numpy / plain torch fp32) of the reference's algorithm for the hot path, each
function citing the reference file:line it follows.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package, and only as the
checker or the timed CPU baseline — never from streamchat_b200/ (the product path has no CPU
fallback and fails loudly without its CUDA library).

Pinning status:
  * kmeans.weighted_kmeans      pinned against the reference function itself (AST-extracted
                                from /root/reference/utiles.py by tests/golden/make_golden.py)
  * kmeans.torch_kmeans_forward pinned against the vendored torch_kmeans package run here
  * memory.*                    pinned against the reference functions run here
  * retrieval.cos_sim           restates sentence-transformers==2.2.2 util.cos_sim (dependency
                                absent from /root/reference and from this image): PARITY UNPINNED
                                beyond the reference's own call sites (utiles.py:17,732,764)
  * vit / qwen2                 pinned against the reference's longva classes + HF transformers
                                5.5.0 (pinned upstream: 4.37.2) run here on tiny configs
  * memory_bank.*               restates langchain 0.0.142 FAISS + faiss-cpu IndexFlatL2 + the patched
                                neighbour expansion (local_doc_qa.py:135-178); packages absent: pinned only
                                against an independent brute force of the published IndexFlatL2 definition
                                (tests/golden/flat_l2_ref.npz) — PARITY UNPINNED against the faiss binary
  * text_encoder.*              pinned against HF BertModel run here (BERT and MiniLM-shaped tiny configs)
  * preprocess.*                pinned bit-exactly against Pillow + HF CLIPImageProcessor run here
The reference ships no tests, golden vectors or fixtures of its own.

oracle/_ref/ (git-ignored; staged by oracle/stage_ref.py from /root/reference during build()) is an UNMODIFIED copy
of the reference's hot-path sources.  oracle/ref_runner.py runs it: the reference's own LlavaQwenForCausalLM /
weighted_kmeans_feature / torch_kmeans on the host cores for bench.py's cpu_baseline and --impl reference legs
(kind "reference"), and tests/test_gpu_dropin_driver.py AST-extracts the driver's own functions from it.
"""
