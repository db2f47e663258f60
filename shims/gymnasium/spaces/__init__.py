from mobrob_b200.spaces import Box  # noqa: F401
