"""``RecordVideo`` is imported unconditionally by examples/control.py:5 and only used with --video-path;
rendering is out of scope, so constructing it fails loudly."""
from mobrob_b200.envs.wrapper import TimeLimit  # noqa: F401


class RecordVideo:
    def __init__(self, env, video_folder, *args, **kwargs):
        raise NotImplementedError("video recording needs MuJoCo's renderer; mobrob_b200 has no renderer "
                                  "(run examples/control.py with --no-gui and without --video-path)")
