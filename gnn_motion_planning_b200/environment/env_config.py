"""Constants of reference environment/env_config.py:3-5."""
import numpy as np

RRT_EPS = 5e-2
STICK_LENGTH = 1.5 * 2 / 15
LIMITS = np.array([1., 1., 8. * RRT_EPS])
