"""CPU oracle for the short-memory sampling helpers (TEST INFRASTRUCTURE).

Follows utiles.py:251-262 of the reference.
"""
from __future__ import annotations

import numpy as np


def calculate_forgetting_probabilities(length: int, tau: float = 10):
    # utiles.py:251-255 — Ebbinghaus curve R_t = exp(-t / tau), normalised; index 0 (the OLDEST
    # frame of the window, hazard H9) gets the largest probability.
    t = np.arange(length)
    r = np.exp(-t / tau)
    return r / r.sum()


def select_data_without_replacement(queue, probabilities, selection_length: int = 10):
    # utiles.py:257-262 — np.random.choice on the GLOBAL numpy RNG (hazard H1), order preserved.
    indices = np.arange(len(queue))
    chosen = np.random.choice(indices, size=selection_length, replace=False, p=probabilities)
    return [queue[i] for i in chosen]
