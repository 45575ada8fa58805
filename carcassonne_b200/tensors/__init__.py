"""Contraction recipes (reference carcassonne/tensors/)."""
