import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O

    O.build()
    return O


EDGE_MODES = {
    "movement_mode": "xy",
    "control_mode": "TCP_velocity_control",
    "noise_mode": "rand_height",
    "observation_mode": "tactile",
    "reward_mode": "dense",
    "arm_type": "ur5",
    "tactile_sensor_name": "tactip",
}


@pytest.fixture(scope="session")
def edge_modes():
    return dict(EDGE_MODES)
