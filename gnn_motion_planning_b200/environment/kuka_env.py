"""Host-side mirrors of reference ``environment/kuka_env.py`` (KukaEnv: 7-DoF, or 13-DoF with model_3.urdf) and
``environment/kuka_2arm_env.py`` (Kuka2Env: two 7-DoF arms, 14-D).

Same duck-typed env protocol as the reference classes.  The reference answers collision queries with PyBullet
contact generation, which cannot be installed or pinned here; these classes answer them
with this repository's geometric model on the GPU (``csrc/arm.cu``: FK + inscribed link spheres vs obstacle
AABBs, arm-vs-arm for 14-D).  PyBullet parity is therefore UNPINNED; everything around the contact query
(joint limits, counters, interpolation, sampling RNG stream) follows the reference.  No CPU fallback.
"""
import pickle

import numpy as np
import torch

from .. import collision


class KukaEnv:
    RRT_EPS = 0.5
    voxel_r = 0.1
    _MODEL_OF_FILE = {"kuka_iiwa/model_0.urdf": collision.ARM_KUKA7, "kuka_iiwa/model_3.urdf": collision.ARM_KUKA13}

    def __init__(self, GUI=False, kuka_file="kuka_iiwa/model_0.urdf", map_file='maze_files/kukas_7_3000.pkl', device=None,
                 problems=None):
        if GUI:
            raise NotImplementedError("no GUI on the B200 path")
        self.dim = 3
        self.kuka_file = kuka_file
        self._model = self._MODEL_OF_FILE[kuka_file]
        self.collision_check_count = 0
        self.collision_time = 0
        self.maps = {}
        self.episode_i = 0
        self.collision_point = None
        if problems is None:
            with open(map_file, 'rb') as f:
                problems = pickle.load(f)
        self.problems = problems
        self.device = torch.device(device if device is not None else "cuda")
        self.config_dim, lo, hi = collision.arm_model_info(self._model)
        self.pose_range = [(float(a), float(b)) for a, b in zip(lo, hi)]
        self.bound = np.array(self.pose_range).T.reshape(-1)
        self.kukaEndEffectorIndex = self.config_dim - 1
        self.order = list(range(len(self.problems)))
        self._boxes, self._box_ptr = collision.pack_boxes([p[0] for p in self.problems], self.device)
        self._problem = 0
        self.k = 0

    def __str__(self):
        return 'kuka' + str(self.config_dim)

    def init_new_problem(self, index=None):
        self.index = self.episode_i if index is None else index
        index = self.index
        obstacles, start, goal, path = self.problems[index]
        self._problem = index
        self.episode_i += 1
        self.episode_i = self.episode_i % len(self.order)
        self.collision_check_count = 0
        self.collision_time = 0
        self.collision_point = None
        self.obstacles = obstacles
        self.init_state = start
        self.goal_state = goal
        self.path = path
        return self.get_problem()

    def get_problem(self):
        return {"init_state": self.init_state, "goal_state": self.goal_state, "obstacles": self.obstacles}

    def set_random_init_goal(self):
        while True:
            points = self.sample_n_points(n=2)
            init, goal = points[0], points[1]
            if np.sum(np.abs(init - goal)) != 0:
                break
        self.init_state, self.goal_state = init, goal

    # ------------------------------------------------------------------ sampling (kuka_env.py:194-222)
    def uniform_sample(self, n=1):
        pr = np.array(self.pose_range)
        sample = np.random.uniform(pr[:, 0], pr[:, 1], size=(n, self.config_dim))
        return sample.reshape(-1) if n == 1 else sample

    def sample_n_points(self, n, need_negative=False):
        """Rejection sampling; consumes the global NumPy RNG exactly like the reference's one-draw-at-a-time loop."""
        pr = np.array(self.pose_range)
        samples, negative = [], []
        need = n
        while need > 0:
            chunk = max(64, int(need * 2.5))
            state = np.random.get_state()
            draws = np.random.uniform(pr[:, 0], pr[:, 1], size=(chunk, self.config_dim))
            free = self.state_fp_batch(draws, count=False)
            idx = np.flatnonzero(free)
            used = chunk if len(idx) < need else int(idx[need - 1]) + 1
            if used < chunk:
                np.random.set_state(state)
                draws = np.random.uniform(pr[:, 0], pr[:, 1], size=(used, self.config_dim))
                free = free[:used]
            self.collision_check_count += used
            for s, f in zip(draws, free):
                (samples if f else negative).append(s)
            need = n - len(samples)
        return (samples, negative) if need_negative else samples

    # ------------------------------------------------------------------ metric helpers (kuka_env.py:224-249)
    def distance(self, from_state, to_state):
        pr = np.array(self.pose_range)
        to_state = np.minimum(np.maximum(to_state, pr[:, 0]), pr[:, 1])
        diff = np.abs(to_state - from_state)
        return np.sqrt(np.sum(diff ** 2, axis=-1))

    def interpolate(self, from_state, to_state, ratio):
        pr = np.array(self.pose_range)
        new_state = from_state + (to_state - from_state) * ratio
        return np.minimum(np.maximum(new_state, pr[:, 0]), pr[:, 1])

    def in_goal_region(self, state):
        return bool(self.distance(state, self.goal_state) < self.RRT_EPS and self._state_fp(state))

    def step(self, state, action=None, new_state=None, check_collision=True):
        pr = np.array(self.pose_range)
        if action is not None:
            new_state = state + action
        new_state = np.minimum(np.maximum(new_state, pr[:, 0]), pr[:, 1])
        action = new_state - state
        if not check_collision:
            return new_state, action
        done = False
        no_collision = self._edge_fp(state, new_state)
        if no_collision and self.in_goal_region(new_state):
            done = True
        return new_state, action, no_collision, done

    # ------------------------------------------------------------------ collision
    def _to_dev(self, x):
        x = np.ascontiguousarray(x)
        if x.dtype not in (np.float32, np.float64):
            x = x.astype(np.float64)
        return torch.from_numpy(x.reshape(-1, self.config_dim)).to(self.device)

    def state_fp_batch(self, states, count=True):
        s = self._to_dev(states)
        prob = torch.full((s.shape[0],), self._problem, dtype=torch.int32, device=self.device)
        free, counted = collision.arm_state_fp(self._model, s, self._boxes, self._box_ptr, prob, want_counted=True)
        if count:
            self.collision_check_count += int(counted.sum())
        return free.cpu().numpy().astype(bool)

    def edge_fp_batch(self, a, b, count=True):
        a = np.asarray(a)
        a_d = self._to_dev(a)
        b_d = self._to_dev(np.asarray(b, dtype=a_d.cpu().numpy().dtype))
        prob = torch.full((a_d.shape[0],), self._problem, dtype=torch.int32, device=self.device)
        free, checks = collision.arm_edge_fp(self._model, a_d, b_d, self._boxes, self._box_ptr, prob, rrt_eps=self.RRT_EPS,
                                             want_checks=True)
        self._last_checks = checks.cpu().numpy()
        if count:
            self.collision_check_count += int(self._last_checks.sum())
        return free.cpu().numpy().astype(bool)

    def _valid_state(self, state):
        pr = np.array(self.pose_range)
        state = np.asarray(state)
        return bool((state >= pr[:, 0]).all() and (state <= pr[:, 1]).all())

    def _point_in_free_space(self, state):
        state = np.asarray(state)
        free = bool(self.state_fp_batch(state.reshape(1, -1))[0])
        if not free and self._valid_state(state):
            self.collision_point = state
        return free

    def _state_fp(self, state):
        return self._point_in_free_space(state)

    def _edge_fp(self, state, new_state):
        state, new_state = np.asarray(state), np.asarray(new_state)
        assert state.size == new_state.size
        self.k = 0
        if new_state.dtype != state.dtype:
            common = np.result_type(state.dtype, new_state.dtype)
            state, new_state = state.astype(common), new_state.astype(common)
        return bool(self.edge_fp_batch(state.reshape(1, -1), new_state.reshape(1, -1))[0])


class Kuka2Env(KukaEnv):
    """reference environment/kuka_2arm_env.py: two iiwa arms based at x = -0.5 / +0.5, 14-D configuration."""
    kukaEndEffectorIndex = 6

    def __init__(self, GUI=False, kuka_file="kuka_iiwa/model.urdf", map_file='maze_files/kukas_14_3000.pkl', device=None,
                 problems=None):
        # "kuka_iiwa/model.urdf" lives in the pybullet_data package, not in the reference tree;
        # the B200 model uses the reference's own model_0.urdf chain for both arms.
        self._MODEL_OF_FILE = {kuka_file: collision.ARM_KUKA14}
        super().__init__(GUI=GUI, kuka_file=kuka_file, map_file=map_file, device=device, problems=problems)
        self.kukaEndEffectorIndex = 6


class UR5Env(KukaEnv):
    """reference environment/ur5_env.py: 6-DoF UR5 on a ground plane, self collision enabled, box obstacles."""
    RRT_EPS = 0.1
    voxel_r = 0.1

    def __init__(self, GUI=False, map_file='maze_files/ur5s_6_3000.pkl', device=None, problems=None):
        self._MODEL_OF_FILE = {"ur5/ur5.urdf": collision.ARM_UR5}
        super().__init__(GUI=GUI, kuka_file="ur5/ur5.urdf", map_file=map_file, device=device, problems=problems)

    def __str__(self):
        return 'ur5'


class SnakeEnv(KukaEnv):
    """reference environment/snake_env.py: 7-D planar snake (base x, y + 5 angles, of which config[3] drives both the
    base yaw and joint 3 and config[6] is unused -- snake_env.py:124-128) in a maze of boxes, self collision enabled."""
    RRT_EPS = 0.1
    voxel_r = 0.1
    height = 0.5

    def __init__(self, map_file='maze_files/snakes_15_2_3000.npz', GUI=False, device=None, maps=None, init_states=None,
                 goal_states=None):
        if GUI:
            raise NotImplementedError("no GUI on the B200 path")
        if maps is None:
            with np.load(map_file) as f:
                maps, init_states, goal_states = f['maps'], f['init_states'], f['goal_states']
        self.maps, self.init_states, self.goal_states = maps, init_states, goal_states
        self.dim = 2
        self._model = collision.ARM_SNAKE7
        self.collision_check_count = 0
        self.size = self.maps.shape[0]
        self.width = self.maps.shape[1]
        self.episode_i = 0
        self.order = list(range(self.size))
        self.collision_point = None
        self.device = torch.device(device if device is not None else "cuda")
        self.config_dim, lo, hi = collision.arm_model_info(self._model)
        self.pose_range = [(float(a), float(b)) for a, b in zip(lo, hi)]       # snake_env.py:55
        self.bound = np.array(self.pose_range).T.reshape(-1)
        self._boxes, self._box_ptr = collision.pack_boxes(collision.snake_obstacles(self.maps), self.device)
        self._problem = 0
        self.k = 0

    def __str__(self):
        return 'snake' + str(self.config_dim)

    def init_new_problem(self, index=None):
        if index is None:
            index = self.episode_i
        self.episode_i += 1
        self.episode_i = self.episode_i % len(self.order)
        self.collision_check_count = 0
        self._problem = index
        self.map = self.maps[index]
        occ = np.argwhere(self.map == 1)
        self.obstacles = occ / self.map.shape[0] - 0.5          # snake_env.py:148-152
        self.collision_point = None
        self.init_state = self.init_states[index]
        self.goal_state = self.goal_states[index]
        return self.get_problem()

    def get_problem(self):
        return {"map": self.map, "init_state": self.init_state, "goal_state": self.goal_state}

    def get_robot_points(self, config):
        return np.array(config[:2])
