from streamchat_b200.kmeans import ClusterResult, KMeans, LpDistance  # noqa: F401

__all__ = ["KMeans", "LpDistance", "ClusterResult"]
