"""Random, filter-normalised directions in weight space (reference: VisTools/directions.py:66-115).

d = randn_like(w);  d *= ||w|| / (||d|| + 1e-10);  parameters with dim <= 1 get a zero direction.
Tiny O(#parameters) set-up work done once per sweep with torch on the model's device; nothing here is on the hot path.
"""
import torch


def get_weights(model):
    """:83-85"""
    return [p.data for p in model.parameters()]


def get_random_weights(weights):
    """:93-98 (drawn on the weights' own device)"""
    return [torch.randn(w.size(), device=w.device) for w in weights]


def normalize_direction(direction, weights):
    """:101-103"""
    for d, w in zip(direction, weights):
        d.mul_(w.norm() / (d.norm() + 1e-10))


def normalize_directions_for_weights(direction, weights):
    """:106-111"""
    assert len(direction) == len(weights)
    for d, w in zip(direction, weights):
        if d.dim() <= 1:
            d.fill_(0)
        normalize_direction(d, w)


def create_random_direction(model):
    """:74-80"""
    weights = get_weights(model)
    direction = get_random_weights(weights)
    normalize_directions_for_weights(direction, weights)
    return direction


def create_random_directions(model):
    """:66-71"""
    return [create_random_direction(model), create_random_direction(model)]
