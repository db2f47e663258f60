"""CPU oracle for the mobrob goal-conditioned PPO hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mobrob_b200/`` may import this
package: it is the checker that ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` compare the CUDA
path against.  Each module cites the reference file:line it restates
(paths relative to the reference checkout).

Pinning status: the reference has no tests and its
arithmetic lives in un-vendored binaries (MuJoCo 2.1.0 via mujoco-py 2.1.2.14,
stable-baselines3 2.0.0, gymnasium 0.28.1) that cannot be installed here, so
the oracle is pinned on the artefacts the reference ships:

* KAT-1  point ``_last_obs`` (real MuJoCo sensor rows) -> ``point_oracle``
         reproduces the stored accelerometer from the stored velocity/gyro.
* KAT-2  ``policy.pth`` forward on ``_last_obs`` -> ``sb3_oracle`` mu / V.
* KAT-3  car ``_last_obs`` layout invariants.
* numpy itself for the RNG streams (PCG64 / MT19937 are numpy's own).

The Euler-with-implicit-damping *integrator form* of ``mj_step`` and the car
contact model are NOT pinned by any artefact ("parity unpinned" for those two
items).
"""
