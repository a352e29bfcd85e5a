"""CPU restatement of the reference Robot's per-step hooks.  TEST INFRASTRUCTURE.

Follows /root/reference/robot.py:
  get_next_action_training / _testing, residual_action, generate_noise   robot.py:541-642
  process_transition, compute_reward, check_if_stuck                    robot.py:645-675, 727-762, 509-538
  get_next_action_type, reset                                           robot.py:443-506
Pinned against tests/golden/robot_golden.npz (the unmodified reference on a deterministic trace).
"""
import numpy as np

from . import td3_oracle as to

NUM_DEMO, PATH_LENGTH, PATH_INCREASE = 3, 50, 20
INITIAL_NOISE, NOISE_DECAY = 1, 0.75
STUCK_THRESHOLD, STUCK_STEPS, STUCK_PENALTY, GOAL_REWARD, DEMO_PROXIMITY_FACTOR = 2, 5, 50, 50, 10
MAX_ACTION, GOAL_RADIUS = 5, 5


class RobotOracle:
    def __init__(self, goal_state, actor_flat=None, hidden=200, layers=3):
        self.goal_state = np.asarray(goal_state, dtype=np.float64)
        self.actor_flat, self.hidden, self.layers = actor_flat, hidden, layers
        self.demonstration_states = []
        self.num_episodes = 0
        self.current_noise_scale = INITIAL_NOISE
        self.path_length = PATH_LENGTH
        self.plan_index = 0
        self.previous_states = []
        self.goal_reached = self.demo_flag = self.stuck_flag = False
        self.updates = 0

    # robot.py:541-595; `unit_noise` = the two standard normals np.random.normal would draw (None: testing)
    def act(self, state, unit_noise=None):
        baseline = np.asarray(state, dtype=np.float64) - self.goal_state
        residual, _ = to.actor_forward(self.actor_flat, baseline.astype(np.float32)[None], self.hidden, self.layers)
        corrected = baseline + residual[0]
        if unit_noise is not None:
            corrected = corrected + (0 + self.current_noise_scale * MAX_ACTION * np.asarray(unit_noise, dtype=np.float64))
        return np.clip(corrected, -MAX_ACTION, MAX_ACTION)

    # robot.py:727-762
    def compute_reward(self, path):
        goal_distance_reward = -np.linalg.norm(path[-1] - self.goal_state)
        if goal_distance_reward >= -GOAL_RADIUS:
            self.goal_reached = True
            return GOAL_REWARD
        if not len(self.demonstration_states):
            return goal_distance_reward
        demos = np.asarray(self.demonstration_states, dtype=np.float64)
        mins = [np.sqrt(((demos - step) ** 2).sum(axis=1)).min() for step in path]
        prox = -np.mean(mins) if self.demo_flag else 0
        return goal_distance_reward + DEMO_PROXIMITY_FACTOR * prox

    # robot.py:509-538
    def check_if_stuck(self, state):
        stuck = False
        if len(self.previous_states) >= STUCK_STEPS:
            diffs = [np.linalg.norm(np.array(state) - np.array(p)) for p in self.previous_states[-STUCK_STEPS:]]
            if all(d < STUCK_THRESHOLD for d in diffs):
                stuck = True
                self.previous_states.clear()
            else:
                self.previous_states.pop(0)
        self.previous_states.append(state)
        return stuck

    # robot.py:645-675 ; returns (reward, done) - the row the reference pushes
    def process_transition(self, state, action, next_state):
        reward = self.compute_reward([next_state])
        if self.check_if_stuck(state):
            self.stuck_flag = True
            reward -= STUCK_PENALTY
        done = self.plan_index == (self.path_length - 1)
        return reward, done

    # robot.py:443-506
    def get_next_action_type(self):
        action_type = "step"
        if self.num_episodes <= NUM_DEMO and not self.demo_flag:
            self.num_episodes += 1
            action_type = "demo"
        if self.num_episodes > NUM_DEMO and not self.demo_flag:
            self.demo_flag = True
            self.num_episodes += 1
            action_type = "reset"
        if self.plan_index == (self.path_length - 1) or self.goal_reached or self.stuck_flag:
            self.num_episodes += 1
            self.plan_index = 0
            self.goal_reached = False
            self.stuck_flag = False
            self.current_noise_scale *= NOISE_DECAY
            self.path_length += PATH_INCREASE
            self.updates += 1
            action_type = "reset"
        else:
            self.plan_index += 1
        return action_type


# ---- the candidate-list rule of the nearest-demonstration search (csrc/rtd3_robot.cu: demo_lists_kernel), restated -------------
# robot.py:753 takes cdist(next_state, demonstration_states).min() over ALL states.  The device evaluates, for a query in the
# 1 x 1 cell c, only the states of c's candidate list.  The rule that builds the lists is restated here so that its claim - the
# minimum over the list equals the minimum over all states, for every point of the cell - can be checked on the CPU against the
# reference's own expression.
def demo_candidate_lists(points, grid=100, cell=1.0, keep=1.0 + 1e-9):
    """Per cell (x-major index cx * grid + cy): indices of the states kept.  A state p is dropped when, for one of five anchor
    states p* (the states nearest to the cell's four corners and to its centre), p* is closer than p at ALL four corners:
    |q-p|^2 - |q-p*|^2 is linear in q, so p* is then closer everywhere in the (convex) cell."""
    P = np.asarray(points, dtype=np.float64)
    lists = []
    for cx in range(grid):
        x0, x1 = cx * cell, (cx + 1) * cell
        dx0, dx1 = (x0 - P[:, 0]) ** 2, (x1 - P[:, 0]) ** 2
        for cy in range(grid):
            y0, y1 = cy * cell, (cy + 1) * cell
            dy0, dy1 = (y0 - P[:, 1]) ** 2, (y1 - P[:, 1]) ** 2
            d = np.stack([dx0 + dy0, dx0 + dy1, dx1 + dy0, dx1 + dy1])                  # [4 corners, M]
            dc = (x0 + 0.5 * cell - P[:, 0]) ** 2 + (y0 + 0.5 * cell - P[:, 1]) ** 2
            stars = [int(np.argmin(d[k])) for k in range(4)] + [int(np.argmin(dc))]
            ok = np.ones(P.shape[0], dtype=bool)
            for st in stars:
                ok &= (d <= d[:, st:st + 1] * keep).any(axis=0)
            lists.append(np.nonzero(ok)[0])
    return lists


def nearest_demo_distance(query, points, lists=None, grid=100, cell=1.0):
    """min_j ||query - points[j]|| as robot.py:753 evaluates it (float64), over the query's candidate list when `lists` is given."""
    P = np.asarray(points, dtype=np.float64)
    q = np.asarray(query, dtype=np.float64)
    if lists is not None and 0 <= q[0] < grid * cell and 0 <= q[1] < grid * cell:
        P = P[lists[int(q[0] / cell) * grid + int(q[1] / cell)]]
    return np.sqrt(((P - q) ** 2).sum(axis=1)).min()
