from . import video_recorder  # noqa: F401
