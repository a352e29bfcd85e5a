"""CPU restatement of the reference's demonstration path.  TEST INFRASTRUCTURE.

Follows /root/reference:
  Environment.get_demonstration (cross-entropy-method planner)     environment.py:140-179
  Robot.augment_demonstration_data                                 robot.py:771-823
  Robot.process_demonstration (replay rows)                        robot.py:679-718
on an explicit numpy-legacy stream (oracle/mt19937.py).  Pinned against tests/golden/demo_golden.npz, produced by the unmodified
reference with its np.random.choice / np.random.normal results recorded (tests/test_oracle_demo.py).
"""
import numpy as np

from . import env_oracle as eo
from .mt19937 import LegacyMT19937

CEM_ITERATIONS, CEM_PATHS, CEM_PATH_LENGTH, CEM_ELITES = 4, 100, 200, 10      # constants.py:37-40
MAX_ACTION = 5
NUM_AUGMENTS, AUG_NOISE, AUG_INTERPOLATION = 3, 2.5, 5                         # robot.py:23-25
GOAL_REWARD, GOAL_RADIUS = 50, 5


def draw_iteration_actions(rng: LegacyMT19937, iteration, mean=None, std=None, paths=CEM_PATHS, steps=CEM_PATH_LENGTH, raw=False):
    """planning_actions[iteration] (float32 [P,T,2]) in the reference's draw order (path, step, component): iteration 0
    np.random.choice([-5, 5], 2) = one 32-bit word & 1 per component; later np.random.normal(mean[step], std[step]) =
    loc + scale * legacy_gauss() per component in float64, stored as float32 (environment.py:154-160).  raw=True: the float64 draws
    themselves (what the reference feeds to `dynamics` before storing the float32 copy)."""
    out = np.zeros((paths, steps, 2), dtype=np.float64 if raw else np.float32)
    for p in range(paths):
        for t in range(steps):
            for c in range(2):
                if iteration == 0:
                    out[p, t, c] = MAX_ACTION if (rng.random_uint32() & 1) else -MAX_ACTION
                else:
                    out[p, t, c] = float(mean[t, c]) + float(std[t, c]) * rng.gauss()
    return out


def refit(actions, rewards, elites=CEM_ELITES):
    """environment.py:168-173: indices of the best paths (ascending), float32 mean / std of their actions, argmax."""
    order = np.argsort(rewards.copy())
    idx = order[-elites:]
    return idx, np.mean(actions[idx], axis=0), np.std(actions[idx], axis=0), int(np.argmax(rewards))


def plan(rng: LegacyMT19937, speed, angle, goal, region, iterations=CEM_ITERATIONS, paths=CEM_PATHS, steps=CEM_PATH_LENGTH, elites=CEM_ELITES):
    """The whole planner on the oracle world (float64 states, as the reference carries them).  Returns (states, actions) float32."""
    start = eo.random_init_state(rng, region)
    mean = std = None
    for it in range(iterations):
        draws = draw_iteration_actions(rng, it, mean, std, paths, steps, raw=True)
        actions = draws.astype(np.float32)                             # planning_actions is a float32 array
        paths_arr = np.zeros((paths, steps + 1, 2), dtype=np.float32)
        rewards = np.zeros(paths)
        for p in range(paths):
            s = np.copy(start)
            paths_arr[p, 0] = s
            for t in range(steps):
                s = eo.dynamics_scalar(speed, angle, s, draws[p, t])     # the float64 draw, environment.py:161
                paths_arr[p, t + 1] = s
            rewards[p] = -np.linalg.norm(paths_arr[p, -1] - goal)
        _, mean, std, best = refit(actions, rewards, elites)
    return paths_arr[best, 0:steps], actions[best]


def augment(rng: LegacyMT19937, demonstration_states, demonstration_actions, noise_level=AUG_NOISE, interpolation_steps=AUG_INTERPOLATION,
            num_augmentations=NUM_AUGMENTS):
    """robot.py:771-823: the augmented states appended to demonstration_states (float64 [A,2]), with the reference's draws - the
    action noise included (np.random.normal(0, s, shape (2,)) = 0 + s * legacy_gauss() per component)."""
    S, A = np.asarray(demonstration_states), np.asarray(demonstration_actions)
    noise = lambda: np.array([0.0 + noise_level * rng.gauss(), 0.0 + noise_level * rng.gauss()])
    out = []
    for _ in range(num_augmentations):
        for i in range(len(S) - 1):
            cur, nxt = S[i], S[i + 1]
            for step in range(1, interpolation_steps + 1):
                fraction = step / float(interpolation_steps + 1)
                synthetic = cur + fraction * (nxt - cur)              # float32 (a Python float is weak against the float32 arrays)
                out.append(synthetic + noise())
                noise()                                               # the action's noise
            out.append(cur + noise())
            noise()
        out.append(S[-1] + noise())
        noise()
    return np.asarray(out, dtype=np.float64)


def demonstration_rows(demonstration_states, demonstration_actions, goal, demo_flag=False, demo_set=None):
    """The T-1 replay rows of robot.py:700-716: (state, action, reward, next_state, done)."""
    S, A = np.asarray(demonstration_states), np.asarray(demonstration_actions)
    rows = []
    for i in range(len(S) - 1):
        gd = -np.linalg.norm(S[i + 1] - goal)
        if gd >= -GOAL_RADIUS:
            reward = GOAL_REWARD
        else:
            prox = 0
            if demo_flag and demo_set is not None and len(demo_set):
                prox = -np.sqrt(((np.asarray(demo_set, np.float64) - S[i + 1]) ** 2).sum(axis=1)).min()
            reward = gd + 10 * prox
        rows.append((S[i], A[i], reward, S[i + 1], i == len(S) - 2))
    return rows
