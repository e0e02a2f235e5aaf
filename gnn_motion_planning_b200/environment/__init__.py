"""Environment mirrors (reference ``environment/``): same duck-typed env protocol, collision checks on the GPU."""
from .kuka_env import Kuka2Env, KukaEnv, SnakeEnv, UR5Env  # noqa: F401
from .maze_env import MazeEnv  # noqa: F401

strs = ['maze2', 'kuka7', 'snake7', 'kuka13', 'ur5', 'kuka14']
