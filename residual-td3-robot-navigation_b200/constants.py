"""Constants of the reference, restated with the same names and values (/root/reference/constants.py:6-53)."""

WORLD_SIZE = 100                 # constants.py:6
ROBOT_RADIUS = 1.5               # constants.py:9
GOAL_RADIUS = 2                  # constants.py:13
INIT_REGION_SIZE = 25            # constants.py:25
UPDATE_RATE = 10                 # constants.py:31
ROBOT_MAX_ACTION = 5             # constants.py:34
DEMOS_CEM_NUM_ITERATIONS = 4     # constants.py:37
DEMOS_CEM_NUM_PATHS = 100        # constants.py:38
DEMOS_CEM_PATH_LENGTH = 200      # constants.py:39
DEMOS_CEM_NUM_ELITES = 10        # constants.py:40
STARTING_MONEY = 100             # constants.py:43
COST_PER_STEP = 0.01             # constants.py:44
COST_PER_CPU_SECOND = 0.03       # constants.py:45
COST_PER_DEMO = 20               # constants.py:46
COST_PER_RESET = 5               # constants.py:47
TEST_DISTANCE_THRESHOLD = 5      # constants.py:50
TEST_TIMEOUT = 100               # constants.py:53
