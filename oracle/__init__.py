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
The reference ships no tests, golden vectors or fixtures of its own.
"""
