"""``explain.cf_example`` of the reference (explain/cf_example.py:8-170) served by the device path of ``icf_b200.explain``:
same class names, constructor arguments and ``explain`` signatures."""
import torch

from icf_b200.explain import DeepCounterfactualExplainer, HingeLossCFExplainer, max_excluding, mse  # noqa: F401


def hinge(true, pred):
    """cf_example.py:8-9."""
    return torch.relu(1 - true * pred)
