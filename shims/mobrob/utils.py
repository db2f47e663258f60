"""Name shim for src/mobrob/utils.py: DATA_DIR, PROJ_DIR, load_policy; the PyBullet recorder is a stub
(drone / turtlebot3 are out of scope)."""
from mobrob_b200.utils import DATA_DIR, PROJ_DIR, load_policy  # noqa: F401


class BulletVideoRecorder:
    def __init__(self, client_id, store_path):
        raise NotImplementedError("PyBullet environments (drone, turtlebot3) are outside mobrob_b200's scope")
