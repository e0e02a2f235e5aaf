"""Host-side mirror of reference ``environment/maze_env.py`` (2-D point robot in a 15x15 grid).

Same duck-typed env protocol as the reference ``MazeEnv``: attributes
``dim, config_dim, bound, RRT_EPS, init_state, goal_state, obstacles, collision_check_count, map,
maps, order, episode_i, k`` and methods ``init_new_problem, sample_n_points, uniform_sample,
distance, interpolate, in_goal_region, step, _valid_state, _point_in_free_space, _state_fp,
_edge_fp, ...``.  Every collision query -- single (drop-in) or batched -- runs in the CUDA kernels of
``csrc/maze.cu`` and reproduces the reference booleans AND its ``collision_check_count`` /
``self.k`` side effects bit-for-bit.  There is no CPU fallback.

Only ``dim == 2`` is implemented (the 3-D stick variant is not in any BASELINE config and its
smoother weights are missing in the reference itself).
"""
import numpy as np
import torch

from .. import collision
from .env_config import LIMITS, RRT_EPS


class MazeEnv:
    RRT_EPS = RRT_EPS
    voxel_r = 1. / 15

    def __init__(self, dim=2, map_file=None, device=None):
        if dim != 2:
            raise NotImplementedError("only the 2-D maze is implemented on the B200 path (SURVEY.md section 8, a7)")
        self.dim = dim
        self.config_dim = dim
        self.collision_check_count = 0
        if map_file is None:
            map_file = 'maze_files/mazes_15_%d_3000.npz' % dim  # maze_env.py:21
        with np.load(map_file) as f:
            self.maps = f['maps']
            self.init_states = f['init_states']
            self.goal_states = f['goal_states']
        self.size = self.maps.shape[0]
        self.width = self.maps.shape[1]
        if self.width != 15:
            raise NotImplementedError("the maze kernels are specialised for 15x15 maps")
        self.bound = (-1, -1, 1, 1)
        self.order = list(range(self.size))
        self.episode_i = 0
        self.collision_point = None
        self.k = 0
        self.device = torch.device(device if device is not None else "cuda")
        self._maps_d = torch.as_tensor(np.ascontiguousarray(self.maps != 0).astype(np.uint8)).to(self.device)
        self._problem = 0

    def __str__(self):
        return 'maze' + str(self.config_dim)

    # ------------------------------------------------------------------ problem management (maze_env.py:41-83)
    def init_new_problem(self, index=None):
        if index is None:
            index = self.episode_i
        self._problem = self.order[index]
        self.map = self.maps[self._problem]
        self.width = self.map.shape[0]
        self.init_state = self.init_states[self._problem]
        self.goal_state = self.goal_states[self._problem]
        self.episode_i += 1
        self.episode_i = self.episode_i % len(self.order)
        self.collision_point = None
        occ = np.argwhere(self.map == 1)                 # row-major (i, j) order, as the reference double loop
        self.obstacles = occ / self.map.shape[0] - 0.5   # maze_env.py:73-79
        self.collision_check_count = 0
        return self.get_problem()

    def get_problem(self):
        return {"map": self.map, "init_state": self.init_state, "goal_state": self.goal_state}

    def set_random_init_goal(self):
        while True:
            init, goal = self.sample_empty_points(), self.sample_empty_points()
            if np.sum(np.abs(init - goal)) != 0:
                break
        self.init_state, self.goal_state = init, goal

    # ------------------------------------------------------------------ sampling (maze_env.py:85-100, 127-135)
    def uniform_sample(self, n=1):
        sample = np.random.uniform(-LIMITS[:self.dim], LIMITS[:self.dim], (n, self.dim))
        return sample.reshape(-1) if n == 1 else sample

    def sample_n_points(self, n, need_negative=False):
        """Rejection sampling with the reference's exact NumPy RNG stream: the global RandomState is
        advanced by exactly the draws the reference would consume, while the state checks run as GPU batches."""
        samples, negative = [], []
        need = n
        while need > 0:
            chunk = max(64, int(need * 2.5))
            state = np.random.get_state()
            draws = np.random.uniform(-LIMITS[:self.dim], LIMITS[:self.dim], (chunk, self.dim))
            free = self.state_fp_batch(draws, count=False)
            idx = np.flatnonzero(free)
            used = chunk if len(idx) < need else int(idx[need - 1]) + 1
            if used < chunk:  # rewind and consume exactly `used` draws
                np.random.set_state(state)
                draws = np.random.uniform(-LIMITS[:self.dim], LIMITS[:self.dim], (used, self.dim))
                free = free[:used]
            self.collision_check_count += used       # every draw is in range -> one counted lookup each
            for s, f in zip(draws, free):
                (samples if f else negative).append(s)
            need = n - len(samples)
        return (samples, negative) if need_negative else samples

    def sample_empty_points(self):
        return self.sample_n_points(1)[0]

    # ------------------------------------------------------------------ metric helpers (maze_env.py:137-179)
    def distance(self, from_state, to_state):
        diff = np.abs(np.asarray(to_state) - np.asarray(from_state))
        if diff.ndim == 1:
            diff = diff.reshape(1, -1)
        return np.sqrt(np.sum(diff ** 2, axis=-1))

    def interpolate(self, from_state, to_state, ratio):
        return from_state + (to_state - from_state) * ratio

    def in_goal_region(self, state):
        return bool(self.distance(state, self.goal_state) < RRT_EPS and self._state_fp(state))

    def step(self, state, action=None, new_state=None, check_collision=True):
        if action is not None:
            new_state = state + action
        new_state[:2] = new_state[:2].clip(-LIMITS[:-1], LIMITS[:-1])
        action = new_state - state
        if not check_collision:
            return new_state, action
        done = False
        no_collision = self._edge_fp(state, new_state)
        if no_collision and self.in_goal_region(new_state):
            done = True
        return new_state, action, no_collision, done

    def get_robot_points(self, config):
        return [config]

    def free_map(self, w=15):
        return [np.array([1. / w + x * 2. / w - 1., 1. / w + y * 2. / w - 1])
                for x in range(self.map.shape[0]) for y in range(self.map.shape[1]) if self.map[x, y] == 0]

    # ------------------------------------------------------------------ collision: batched (new) ...
    def _to_dev(self, x):
        x = np.ascontiguousarray(x)
        if x.dtype not in (np.float32, np.float64):
            x = x.astype(np.float64)
        return torch.from_numpy(x.reshape(-1, 2)).to(self.device)

    def state_fp_batch(self, states, count=True):
        """bool[n] for n states against the current problem; advances collision_check_count like n _state_fp calls."""
        s = self._to_dev(states)
        prob = torch.full((s.shape[0],), self._problem, dtype=torch.int32, device=self.device)
        free, counted = collision.maze_state_fp(s, self._maps_d, prob, want_counted=True)
        if count:
            self.collision_check_count += int(counted.sum())
        return free.cpu().numpy().astype(bool)

    def edge_fp_batch(self, a, b, count=True):
        """bool[n] for n edges against the current problem; advances collision_check_count like n _edge_fp calls."""
        a_d = self._to_dev(a)
        b_d = self._to_dev(np.asarray(b, dtype=np.asarray(a).dtype if np.asarray(a).dtype in (np.float32, np.float64) else np.float64))
        prob = torch.full((a_d.shape[0],), self._problem, dtype=torch.int32, device=self.device)
        free, checks = collision.maze_edge_fp(a_d, b_d, self._maps_d, prob, want_checks=True)
        checks = checks.cpu().numpy()
        if count:
            self.collision_check_count += int(checks.sum())
        self._last_checks = checks
        return free.cpu().numpy().astype(bool)

    # ------------------------------------------------------------------ ... and the drop-in scalar protocol
    def _valid_state(self, state):
        state = np.asarray(state)
        return bool((state >= -LIMITS[:state.size]).all() and (state <= LIMITS[:state.size]).all())

    def _point_in_free_space(self, state):
        state = np.asarray(state)
        assert state.size == 2
        free = bool(self.state_fp_batch(state.reshape(1, 2))[0])
        if not free and not self._valid_state(state):
            self.collision_point = state
        return free

    def _state_fp(self, state):
        return self._point_in_free_space(state)

    def _edge_fp(self, state, new_state):
        state, new_state = np.asarray(state), np.asarray(new_state)
        assert state.size == new_state.size == 2
        if new_state.dtype != state.dtype:
            common = np.result_type(state.dtype, new_state.dtype)
            state, new_state = state.astype(common), new_state.astype(common)
        free = bool(self.edge_fp_batch(state.reshape(1, 2), new_state.reshape(1, 2))[0])
        # self.k = number of bisection midpoints (maze_env.py:308): lookups minus the two endpoint lookups
        self.k = max(int(self._last_checks[0]) - 2, 0)
        return free
