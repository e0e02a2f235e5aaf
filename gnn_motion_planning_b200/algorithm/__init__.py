"""Callers of the hot path outside eval_gnn: the graph construction the reference's baselines and dataset builder share."""
from .dijkstra import construct_graph  # noqa: F401
