"""Host-side mirror of reference ``str2name.py:11-81``: env-name -> (env, explorer, explorer-weights path, smoother,
smoother-weights path[, data path]).  Same hyper-parameter table; models live on the GPU."""
import numpy as np
import torch

from .environment import Kuka2Env, KukaEnv, MazeEnv, SnakeEnv, UR5Env
from .model import EncoderProcessDecoder
from .model_smoother import ModelSmoother

# name -> (workspace_size, config_size, explorer embed, obs_size, explorer weights, smoother weights, data path)
TABLE = {
    "maze2": (2, 2, 32, 2, "data/weights/weights_maze.pt", "data/weights/smooth_2d_attv3.pt", "data/pkl/maze_prm_4000.pkl"),
    # maze3: the reference's smoother file smooth_3d_attv3.pt is not shipped (only smooth_3d_att.pt, another architecture), so
    # load=True fails there as it does in the reference; the explorer + MazeEnv(dim=3) run with smoother='none'
    "maze3": (2, 3, 32, 2, "data/weights/weights_maze_3.pt", "data/weights/smooth_3d_attv3.pt", "data/pkl/maze_prm_3.pkl"),
    "kuka7": (3, 7, 64, 6, "data/weights/weights_kuka.pt", "data/weights/smooth_7d_attv3.pt", "data/pkl/kuka_prm_4000.pkl"),
    "ur5": (3, 6, 32, 6, "data/weights/weights_ur5.pt", "data/weights/smooth_ur5_attv3.pt", "data/pkl/ur5_prm_3000.pkl"),
    "snake7": (3, 7, 32, 2, "data/weights/weights_snake.pt", "data/weights/smooth_snake_attv3.pt", "data/pkl/snake_prm_3000.pkl"),
    "kuka13": (3, 13, 32, 6, "data/weights/weights_kuka_13.pt", "data/weights/smooth_13d_attv3.pt", "data/pkl/kuka_prm_13.pkl"),
    "kuka14": (3, 14, 32, 6, "data/weights/kuka_14.pt", "data/weights/smooth_14d_attv3.pt", "data/pkl/kuka_prm_14.pkl"),
}


def _make_env(name, **env_kwargs):
    if name == "maze2":
        return MazeEnv(dim=2, **env_kwargs)
    if name == "maze3":
        return MazeEnv(dim=3, **env_kwargs)
    if name == "kuka7":
        return KukaEnv(**env_kwargs)
    if name == "kuka13":
        return KukaEnv(kuka_file="kuka_iiwa/model_3.urdf", map_file=env_kwargs.pop("map_file", "maze_files/kukas_13_3000.pkl"), **env_kwargs)
    if name == "kuka14":
        return Kuka2Env(**env_kwargs)
    if name == "ur5":
        return UR5Env(**env_kwargs)
    if name == "snake7":
        return SnakeEnv(**env_kwargs)
    return None


def str2name(str, get_data=False, use_obstacle=True, load=False, make_env=True, **env_kwargs):
    key = "maze2" if "maze2" in str else str          # reference: `if 'maze2' in str` (str2name.py:12)
    if key not in TABLE:
        raise KeyError("unknown environment %r; known: %s" % (str, sorted(TABLE)))
    ws, c, e, s, explore_path, smooth_path, data_path = TABLE[key]
    device = torch.device("cuda", torch.cuda.current_device())
    env = _make_env(key, **env_kwargs) if make_env else None
    scale = float(np.max(env.bound)) if (key == "ur5" and env is not None) else (2 * np.pi if key == "ur5" else 1.0)  # str2name.py:40
    model_explore = EncoderProcessDecoder(workspace_size=ws, config_size=c, embed_size=e, obs_size=s).to(device)
    model_smooth = ModelSmoother(workspace_size=3, config_size=c, embed_size=128, obs_size=6, scale=scale).to(device)
    if not use_obstacle:
        explore_path = explore_path.replace('.pt', '_pure.pt')   # str2name.py:68-69 (these files are not shipped)
    if load:
        model_explore.load_state_dict(torch.load(explore_path, map_location="cpu"))
        model_smooth.load_state_dict(torch.load(smooth_path, map_location="cpu"))
    if get_data:
        return env, model_explore, explore_path, model_smooth, smooth_path, data_path
    return env, model_explore, explore_path, model_smooth, smooth_path
