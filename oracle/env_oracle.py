"""CPU restatement of the reference ``Environment`` hot path.  TEST INFRASTRUCTURE.

Follows /root/reference/environment.py:
  dynamics                    environment.py:98-119
  step                        environment.py:122-127
  reset / init-state draw     environment.py:130-137
  set_init_and_goal           environment.py:28-56
  compute_reward              environment.py:182-183

Two forms of every function: a scalar one that mirrors the reference statement by
statement (same numpy scalar calls, same dtypes - float64 arithmetic on float32
table values, and under numpy >= 2 / NEP-50 the rotation ``angle*2*pi`` stays
float32), and a vectorised numpy one over n envs that is checked against the scalar
form in tests/test_oracle_env.py.  Pinned against the live reference through
tests/golden/env_golden.npz.
"""
import numpy as np

from .mt19937 import LegacyMT19937

WORLD_SIZE = 100                 # constants.py:6
INIT_REGION_SIZE = 25            # constants.py:25
ROBOT_MAX_ACTION = 5             # constants.py:34
CLIP_HI = WORLD_SIZE - 1.0001    # environment.py:117


def dynamics_scalar(speed_map, angle_map, state, action):
    """environment.py:98-119, statement by statement."""
    action = np.clip(action, -ROBOT_MAX_ACTION, ROBOT_MAX_ACTION)
    action_magnitude = np.linalg.norm(action)
    action_angle = np.arctan2(action[1], action[0])
    cell_x = int(state[0])
    cell_y = int(state[1])
    rotation = angle_map[cell_x, cell_y] * 2 * np.pi
    rotated_action_angle = action_angle + rotation
    speed = speed_map[cell_x, cell_y]
    next_state_x = state[0] + speed * action_magnitude * np.cos(rotated_action_angle)
    next_state_y = state[1] + speed * action_magnitude * np.sin(rotated_action_angle)
    next_state = np.array([next_state_x, next_state_y])
    next_state = np.clip(next_state, 0, CLIP_HI)
    return next_state


def step_scalar(speed_map, angle_map, robot_state, action):
    """environment.py:122-127.  Returns the new robot_state (kept if next is out of [0,100) - only NaN can do that)."""
    next_state = dynamics_scalar(speed_map, angle_map, robot_state, action)
    if 0 <= next_state[0] < WORLD_SIZE and 0 <= next_state[1] < WORLD_SIZE:
        return next_state
    return robot_state


def dynamics_batch(speed_map, angle_map, states, actions):
    """Vectorised environment.py:98-119 over n envs.  states [n,2] (any float), actions [n,2] -> [n,2] float64.

    Same operation order and dtypes as the scalar form: float64 state/action arithmetic,
    float32 rotation (angle*2*pi evaluated in float32), float32 speed promoted at the multiply.
    """
    states = np.asarray(states)
    a = np.clip(np.asarray(actions), -ROBOT_MAX_ACTION, ROBOT_MAX_ACTION)
    mag = np.sqrt(a[:, 0] * a[:, 0] + a[:, 1] * a[:, 1])
    ang = np.arctan2(a[:, 1], a[:, 0])
    cx = states[:, 0].astype(np.int64)      # int() truncates toward zero; states are >= 0
    cy = states[:, 1].astype(np.int64)
    rot = angle_map[cx, cy] * np.float32(2) * np.float32(np.pi)   # float32, as NEP-50 gives the scalar form
    rang = ang + rot
    speed = speed_map[cx, cy]
    nx = states[:, 0] + speed * mag * np.cos(rang)
    ny = states[:, 1] + speed * mag * np.sin(rang)
    nxt = np.stack([nx, ny], axis=1).astype(np.float64)
    return np.clip(nxt, 0, CLIP_HI)


def step_batch(speed_map, angle_map, states, actions):
    """environment.py:122-127 over n envs."""
    nxt = dynamics_batch(speed_map, angle_map, states, actions)
    ok = (0 <= nxt[:, 0]) & (nxt[:, 0] < WORLD_SIZE) & (0 <= nxt[:, 1]) & (nxt[:, 1] < WORLD_SIZE)
    return np.where(ok[:, None], nxt, np.asarray(states, dtype=np.float64))


def set_init_and_goal(rng: LegacyMT19937):
    """environment.py:28-56 on an explicit legacy stream.  Returns (goal[2] f64, region[4] f64 = l,r,b,t, tries)."""
    r = rng.interval(3)                                   # np.random.choice([0,1,2,3])
    span = WORLD_SIZE - INIT_REGION_SIZE
    if r == 0:
        init_left = 0
        init_right = INIT_REGION_SIZE
        init_bottom = rng.uniform(0, span)
        init_top = init_bottom + INIT_REGION_SIZE
    elif r == 1:
        init_left = rng.uniform(0, span)
        init_right = init_left + INIT_REGION_SIZE
        init_bottom = WORLD_SIZE - INIT_REGION_SIZE
        init_top = WORLD_SIZE
    elif r == 2:
        init_left = WORLD_SIZE - INIT_REGION_SIZE
        init_right = WORLD_SIZE
        init_bottom = rng.uniform(0, span)
        init_top = init_bottom + INIT_REGION_SIZE
    else:
        init_left = rng.uniform(0, span)
        init_right = init_left + INIT_REGION_SIZE
        init_bottom = 0
        init_top = INIT_REGION_SIZE
    init_mid = np.array([0.5 * (init_left + init_right), 0.5 * (init_bottom + init_top)])
    distance = 0
    tries = 0
    while distance < 90:
        random_goal = np.array([rng.uniform(5, WORLD_SIZE - 5), rng.uniform(5, WORLD_SIZE - 5)])
        distance = np.linalg.norm(random_goal - init_mid)
        tries += 1
    region = np.array([init_left, init_right, init_bottom, init_top], dtype=np.float64)
    return random_goal, region, tries


def random_init_state(rng: LegacyMT19937, region):
    """environment.py:135-137: uniform([l,b],[r,t],2) - x first, then y."""
    x = rng.uniform(region[0], region[1])
    y = rng.uniform(region[2], region[3])
    return np.array([x, y], dtype=np.float64)


def compute_reward_env(path, goal_state):
    """environment.py:182-183."""
    return -np.linalg.norm(path[-1] - goal_state)


def synthetic_maps(seed=0):
    """The benchmark/parity maps of SURVEY.md section 8(d) config 2 (NOT the reference's Perlin maps - unpinned).

    u = RandomState(seed).rand(100,100) low-pass filtered by three passes of a 5-point
    box blur with edge replication, renormalised to [0,1];
    speed = sigmoid(10*(u-0.5)) (the reference's stretch, environment.py:81-83), angle = u.  Both float32 [x][y].
    """
    u = np.random.RandomState(seed).rand(WORLD_SIZE, WORLD_SIZE)
    for _ in range(3):
        p = np.pad(u, 1, mode="edge")
        u = (p[1:-1, 1:-1] + p[:-2, 1:-1] + p[2:, 1:-1] + p[1:-1, :-2] + p[1:-1, 2:]) / 5.0
    u = (u - u.min()) / (u.max() - u.min())
    u = u.astype(np.float32)
    speed = (1 / (1 + np.exp(-10 * (u - 0.5)))).astype(np.float32)
    return speed, u.copy()
