"""Configuration of the reference (/root/reference/configuration.py:12-26); graphics keys kept for attribute parity."""
from . import constants

WINDOW_SIZE = 550
GRAPHICS_ON = False
MODEL_VISUALISATION_ACTION = [0.5 * constants.ROBOT_MAX_ACTION, 0.5 * constants.ROBOT_MAX_ACTION]
RANDOM_SEED = 1707366464         # configuration.py:26
