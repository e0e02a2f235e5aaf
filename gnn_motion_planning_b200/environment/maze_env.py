"""Host-side mirror of reference ``environment/maze_env.py`` (2-D point robot in a 15x15 grid).

Same duck-typed env protocol as the reference ``MazeEnv``: attributes
``dim, config_dim, bound, RRT_EPS, init_state, goal_state, obstacles, collision_check_count, map,
maps, order, episode_i, k`` and methods ``init_new_problem, sample_n_points, uniform_sample,
distance, interpolate, in_goal_region, step, _valid_state, _point_in_free_space, _state_fp,
_edge_fp, ...``.  Every collision query -- single (drop-in) or batched -- runs in the CUDA kernels of
``csrc/maze.cu`` and reproduces the reference booleans AND its ``collision_check_count`` /
``self.k`` side effects bit-for-bit.  There is no CPU fallback.

``dim == 2`` is the 2-D point robot of the BASELINE configs; ``dim == 3`` (round 2) is the 3-D stick robot
(state = x, y, theta; maze_env.py:245-264, 279-291, 327-347) on ``gmp_maze3_state_fp`` / ``gmp_maze3_edge_fp`` --
its smoother weights are missing in the reference itself, so only ``explore(..., smoother='none')`` runs there.
"""
import numpy as np
import torch

from .. import collision
from .env_config import LIMITS, RRT_EPS, STICK_LENGTH


class MazeEnv:
    RRT_EPS = RRT_EPS
    voxel_r = 1. / 15

    def __init__(self, dim=2, map_file=None, device=None):
        if dim not in (2, 3):
            raise NotImplementedError("MazeEnv: dim must be 2 (point robot) or 3 (stick robot)")
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
        self.bound = (-1, -1, 1, 1) if dim == 2 else (-1, -1, -0.4, 1, 1, 0.4)      # maze_env.py:30-33
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
            checks = self._last_state_checks
            idx = np.flatnonzero(free)
            used = chunk if len(idx) < need else int(idx[need - 1]) + 1
            if used < chunk:  # rewind and consume exactly `used` draws
                np.random.set_state(state)
                draws = np.random.uniform(-LIMITS[:self.dim], LIMITS[:self.dim], (used, self.dim))
                free = free[:used]
            # 2-D: every draw is in range -> one counted lookup each; 3-D: the lookups of each consumed stick check
            self.collision_check_count += int(checks[:used].sum())
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
        if self.dim >= 3:                                    # theta wraps (maze_env.py:145-147)
            diff[:, 2] = np.min((diff[:, 2], np.abs(diff[:, 2] - 2 * LIMITS[2])), axis=0)
        return np.sqrt(np.sum(diff ** 2, axis=-1))

    def interpolate(self, from_state, to_state, ratio):
        diff = to_state - from_state
        if self.dim >= 3:                                    # maze_env.py:154-160
            if np.abs(diff[2]) > LIMITS[2]:
                if diff[2] > 0:
                    diff[2] -= 2 * LIMITS[2]
                else:
                    diff[2] += 2 * LIMITS[2]
        new_state = from_state + diff * ratio
        if self.dim >= 3:                                    # maze_env.py:164-170
            if np.abs(new_state[2]) > LIMITS[2]:
                if new_state[2] > 0:
                    new_state[2] -= 2 * LIMITS[2]
                else:
                    new_state[2] += 2 * LIMITS[2]
        return new_state

    @staticmethod
    def _end_points(coord=None, l=None, center=None, theta=None, a=None, b=None):
        """maze_env.py:245-264 (host helper for plots; the collision kernels compute the end points themselves)."""
        if theta is None:
            theta = coord[2] / LIMITS[2] * np.pi
        orient = np.array([np.cos(theta), np.sin(theta)])
        if l is None:
            l = STICK_LENGTH
        if a is None and b is None:
            if center is None:
                center = np.array(coord[:2])
            a = center - l / 2. * orient
            b = center + l / 2. * orient
        else:
            if a is not None:
                b = a + l * orient
            if b is not None:
                a = b - l * orient
        return a, b

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
    def _to_dev(self, x, width=2):
        x = np.ascontiguousarray(x)
        if x.dtype not in (np.float32, np.float64):
            x = x.astype(np.float64)
        return torch.from_numpy(x.reshape(-1, width)).to(self.device)

    def state_fp_batch(self, states, count=True):
        """bool[n] for n states against the current problem; advances collision_check_count like n _state_fp calls.
        States are 2-D points, or (dim == 3 and 3 columns) sticks."""
        width = 3 if (self.dim == 3 and np.asarray(states).shape[-1] == 3) else 2
        s = self._to_dev(states, width)
        prob = torch.full((s.shape[0],), self._problem, dtype=torch.int32, device=self.device)
        if width == 3:
            free, checks, k = collision.maze3_state_fp(s, self._maps_d, prob)
            checks = checks.cpu().numpy()
            self._last_k = k.cpu().numpy()
        else:
            free, counted = collision.maze_state_fp(s, self._maps_d, prob, want_counted=True)
            checks = counted.cpu().numpy().astype(np.int64)
        self._last_state_checks = checks
        if count:
            self.collision_check_count += int(checks.sum())
        return free.cpu().numpy().astype(bool)

    def edge_fp_batch(self, a, b, count=True):
        """bool[n] for n edges against the current problem; advances collision_check_count like n _edge_fp calls."""
        width = 3 if (self.dim == 3 and np.asarray(a).shape[-1] == 3) else 2
        a_d = self._to_dev(a, width)
        b_d = self._to_dev(np.asarray(b, dtype=np.asarray(a).dtype if np.asarray(a).dtype in (np.float32, np.float64) else np.float64), width)
        prob = torch.full((a_d.shape[0],), self._problem, dtype=torch.int32, device=self.device)
        if width == 3:
            free, checks, k = collision.maze3_edge_fp(a_d, b_d, self._maps_d, prob)
            self._last_k = k.cpu().numpy()
        else:
            free, checks = collision.maze_edge_fp(a_d, b_d, self._maps_d, prob, want_checks=True)
            self._last_k = None
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

    def _stick_in_free_space(self, state):
        state = np.asarray(state)
        assert state.size == 3
        free = bool(self.state_fp_batch(state.reshape(1, 3))[0])
        self.k = int(self._last_k[0])
        if not free and self._valid_state(state):
            self.collision_point = state
        return free

    def _state_fp(self, state):
        state = np.asarray(state)
        assert state.size == 2 or state.size == 3
        return self._point_in_free_space(state) if state.size == 2 else self._stick_in_free_space(state)

    def _edge_fp(self, state, new_state):
        state, new_state = np.asarray(state), np.asarray(new_state)
        assert state.size == new_state.size and state.size in (2, 3)
        if new_state.dtype != state.dtype:
            common = np.result_type(state.dtype, new_state.dtype)
            state, new_state = state.astype(common), new_state.astype(common)
        if state.size == 3:                                  # stick robot (maze_env.py:327-347)
            free = bool(self.edge_fp_batch(state.reshape(1, 3), new_state.reshape(1, 3))[0])
            self.k = int(self._last_k[0])
            return free
        free = bool(self.edge_fp_batch(state.reshape(1, 2), new_state.reshape(1, 2))[0])
        # self.k = number of bisection midpoints (maze_env.py:308): lookups minus the two endpoint lookups
        self.k = max(int(self._last_checks[0]) - 2, 0)
        return free
