"""CPU dry run of the scripts tests/test_gpu_vertical_posctl.py hands to its GPU subprocesses: with tg.make_vec replaced by an
oracle-backed stand-in (tests/_oracle_backed_vec.py) the scripts compare the oracle with itself, which exercises their own logic -
state indices, shapes, tolerances, the "surface really shows" condition - so that a GPU failure is decided by the device code and not by a slip in the test."""
import importlib.util
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _scripts():
    spec = importlib.util.spec_from_file_location("tg_unverified_scripts", os.path.join(HERE, "test_gpu_vertical_posctl.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture()
def fake_make_vec(monkeypatch, oracle):
    sys.path.insert(0, HERE)
    import _oracle_backed_vec as F
    import tactile_gym_b200 as tg

    monkeypatch.setattr(tg, "make_vec", lambda env_id, n, **kw: F.FakeVec(env_id, n, **kw))
    yield
    sys.path.remove(HERE)


@pytest.mark.parametrize("arm,sensor,S,obs", [("mg400", "tactip", 128, "tactile"), ("mg400", "tactip", 64, "oracle")])
def test_vertical_script_logic(fake_make_vec, capsys, arm, sensor, S, obs):
    tu = _scripts()
    code = tu.CHILD % {"root": tu.ROOT, "arm": arm, "sensor": sensor, "S": S, "obs": obs, "render": "True" if obs == "tactile" else "False", "steps": 8}
    exec(compile(code, "vertical-child", "exec"), {"__name__": "child"})
    assert "VERTICAL-OK" in capsys.readouterr().out


def test_mg400_position_control_script_logic(fake_make_vec, capsys):
    tu = _scripts()
    exec(compile(tu.CHILD_POSCTL % {"root": tu.ROOT}, "posctl-child", "exec"), {"__name__": "child"})
    assert "POSCTL-OK" in capsys.readouterr().out
