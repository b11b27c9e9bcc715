class VideoRecorder:
    def __init__(self, *a, **k):
        raise RuntimeError("gym shim: VideoRecorder unavailable")
