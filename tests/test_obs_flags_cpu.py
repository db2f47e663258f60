"""Host logic of the optional observation keys: config parsing and the oracle's row layout."""
import numpy as np
import pytest

from oracle import car_oracle as co, point_oracle as po
from oracle.vec_oracle import GoalVecOracle, extend_obs


def test_robot_config_to_flag_bits():
    from mobrob_b200.vec_env import obs_flags_of

    assert obs_flags_of(None) == 0
    # the reference's own PointEnv / CarEnv configs (wrapper.py:293-317) select the default row
    assert obs_flags_of({"robot_base": "xmls/car.xml", "sensors_obs": ["accelerometer"], "observe_com": False,
                         "observe_goal_comp": True, "box_size": 0.125}) == 0
    assert obs_flags_of({"observe_goal_dist": True}) == 1
    assert obs_flags_of({"observe_qpos": True, "observe_qvel": True, "observe_ctrl": True}) == 2 | 4 | 8
    assert obs_flags_of({"observe_qpos": False}) == 0
    for key, value in (("observe_hazards", True), ("observe_vision", True), ("observe_goal_comp", False), ("observe_com", True)):
        with pytest.raises(NotImplementedError):
            obs_flags_of({key: value})


@pytest.mark.parametrize("body_cls,nq,nv,pre", [(po.PointBody, 3, 3, 3), (co.CarBody, 13, 11, 15)])
def test_oracle_row_is_sorted_key_order(body_cls, nq, nv, pre):
    n = 5
    ora = GoalVecOracle(body_cls(n), seed=1, time_limit=50, terminate_on_goal=True)
    base = ora.reset()
    rng = np.random.default_rng(0)
    for _ in range(3):
        base, *_ = ora.step(rng.uniform(-1.3, 1.3, (n, 2)).astype(np.float32))
    B = base.shape[1]
    every = dict(observe_goal_dist=True, observe_qpos=True, observe_qvel=True, observe_ctrl=True)
    row = extend_obs(ora.body, base, ora.goal, every)
    assert row.dtype == np.float32 and row.shape == (n, B + 1 + nq + nv + 2)
    k = 0
    np.testing.assert_array_equal(row[:, k:k + pre], base[:, :pre]); k += pre
    np.testing.assert_array_equal(row[:, k:k + 2], ora.body.ctrl.astype(np.float32)); k += 2
    np.testing.assert_array_equal(row[:, k:k + 2], base[:, pre:pre + 2]); k += 2
    d = np.linalg.norm(ora.goal.astype(np.float64) - ora.body.pos(), axis=1)
    np.testing.assert_allclose(row[:, k], np.exp(-d), rtol=1e-7); k += 1
    np.testing.assert_array_equal(row[:, k:k + 6], base[:, pre + 2:pre + 8]); k += 6
    np.testing.assert_array_equal(row[:, k:k + nq], ora.body.qpos().astype(np.float32)); k += nq
    np.testing.assert_array_equal(row[:, k:k + nv], ora.body.qvel().astype(np.float32)); k += nv
    np.testing.assert_array_equal(row[:, k:], base[:, B - 3:])
    # no key switched on: the default row, untouched
    assert extend_obs(ora.body, base, ora.goal, {"observe_qpos": False}) is base
    # the vectorised stack returns the extended rows from reset / step / terminal observations
    ext = GoalVecOracle(body_cls(n), seed=1, time_limit=2, terminate_on_goal=True, observe={"observe_ctrl": True})
    assert ext.reset().shape == (n, B + 2)
    o, r, done, info = ext.step(np.ones((n, 2), np.float32))
    o, r, done, info = ext.step(np.ones((n, 2), np.float32))
    assert done.all() and info["terminal_obs"].shape == (n, B + 2)
    np.testing.assert_array_equal(info["terminal_obs"][:, pre:pre + 2], np.ones((n, 2), np.float32))
    np.testing.assert_array_equal(o[:, pre:pre + 2], np.zeros((n, 2), np.float32))   # full reset: ctrl = 0
