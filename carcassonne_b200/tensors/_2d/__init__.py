"""2D corner/side environment recipes (reference carcassonne/tensors/_2d/)."""
