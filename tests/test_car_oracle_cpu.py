"""Car oracle: the warm-started contact sweeps keep an env step a function of (state, action) alone."""
import numpy as np

from oracle import car_oracle as co


def _driven(n, steps, seed):
    rng = np.random.default_rng(seed)
    b = co.CarBody(n)
    for i in range(n):
        b.full_reset(i, rng.uniform(-1, 1, 2), rng.uniform(0, 2 * np.pi))
    for _ in range(steps):
        b.step(np.sign(rng.standard_normal((n, 2))))
    return b, rng


def test_step_and_obs_do_not_depend_on_the_solver_history():
    a, rng = _driven(3, 6, 0)
    b = co.CarBody(3)
    for k in ("p", "quat", "v", "w", "th", "s", "qb", "wb", "ctrl"):
        getattr(b, k)[:] = getattr(a, k)
    assert np.abs(a.last_forces).max() > 0 and not np.any(b.last_forces)   # a carries the previous step's forces, b none
    goal = np.zeros((3, 2), np.float32)
    np.testing.assert_array_equal(a.obs(goal), b.obs(goal))                # observation solves are cold
    act = np.sign(rng.standard_normal((3, 2)))
    a.step(act); b.step(act)
    np.testing.assert_array_equal(a.state_vector(), b.state_vector())      # the first substep of a step is cold
    np.testing.assert_array_equal(a.last_forces, b.last_forces)


def test_warm_sweeps_stay_close_to_cold_ones():
    """Substeps 2..10 run N_SWEEPS_WARM sweeps from the previous forces; one env step differs from the all-cold solve
    by no more than the cold solve's own distance from convergence allows (a few per cent of the contact forces)."""
    a, rng = _driven(4, 5, 1)
    b = co.CarBody(4)
    for k in ("p", "quat", "v", "w", "th", "s", "qb", "wb", "ctrl"):
        getattr(b, k)[:] = getattr(a, k)
    act = np.sign(rng.standard_normal((4, 2)))
    a.step(act)
    b.ctrl[:] = np.clip(act, -1, 1)
    for _ in range(co.FRAME_SKIP):
        b.substep(warm=False)
    assert co.N_SWEEPS_WARM < co.N_SWEEPS
    normal = b.last_forces[:, :, 2].sum(1)
    assert np.all(normal > 0.5 * co.MASS * co.GRAV)                        # the car stands on its contacts
    assert np.abs(a.p - b.p).max() < 2e-3 and np.abs(a.v - b.v).max() < 5e-2
    assert np.abs(a.last_forces[:, :, 2].sum(1) - normal).max() < 0.1 * co.MASS * co.GRAV
