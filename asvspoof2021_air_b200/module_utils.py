"""Glue that exposes an engine's flat buffers as an nn.Module tree with the reference's
state_dict names (parameters/buffers are VIEWS of the flat storage, not copies)."""
import torch.nn as nn


class _Node(nn.Module):
    """Name-only container (e.g. `layer1`, `layer1.0`, `layer1.0.bn1`)."""


def _descend(root, parts):
    node = root
    for part in parts:
        child = node._modules.get(part)
        if child is None:
            child = _Node()
            node.add_module(part, child)
        node = child
    return node


def bind_state(root, ordered_keys, param_views, buffer_views):
    """Register views under dotted reference names, in the reference's state_dict order.

    ordered_keys: list of (key, kind) with kind in {"param", "buffer"}."""
    for key, kind in ordered_keys:
        *parents, leaf = key.split(".")
        node = _descend(root, parents)
        if kind == "param":
            p = nn.Parameter(param_views[key], requires_grad=True)
            if leaf in node._parameters:
                node._parameters[leaf] = p
            else:
                node.register_parameter(leaf, p)
        else:
            if leaf in node._buffers:
                node._buffers[leaf] = buffer_views[key]
            else:
                node.register_buffer(leaf, buffer_views[key])
